// himo_b200/csrc/dec.cuh -- internal interface between dec.cu and deflowpp.cu (same shared object).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "himo_b200.h"

namespace himo {

struct DecGatherArgs {
  const float4* pt4;            // [n] pc0 warped xyz + key
  int n, n_pad;
  int n_frames;                 // before = n_frames*32 channels
  const unsigned* bitmap;       // [F][n_words]
  const int* word_prefix;       // [F][n_words]
  const float* voxel_feats;     // [F][n_max][32]
  int n_words, n_max;
  const float* after;           // [gy*gx][C_after] fp32 NHWC (backbone output)
  int c_after;
  const float* w_off;           // [96][3]
  const float* b_off;           // [96]
  float vx, vy, vz, x_min, y_min, z_min, hx, hy, hz;
  int gx;
  float* h32;                   // [n_pad][192]
  __nv_bfloat16* hx_planes;     // [P][n_pad][288]
  __nv_bfloat16* rhx_planes;    // [P][n_pad][288]
  int planes;
  long long plane_stride;       // n_pad*288
};

int dec_gather(const DecGatherArgs& a, cudaStream_t stream);
int dec_rh(const float* zr, const float* h32, int n_pad, __nv_bfloat16* rhx, int planes, long long ps,
           cudaStream_t stream);
int dec_update(const float* zr, const float* q, float* h32, int n_pad, __nv_bfloat16* hx, int planes,
               long long ps, cudaStream_t stream);
int dec_out(const float* y, int y_stride, const float4* pt4, int n, const float* w2, const float* b2,
            float* flow, cudaStream_t stream);
// csrc/decfused.cu: the whole ConvGRU decoder after the gather in one persistent tcgen05 kernel
int dec_fused(const __nv_bfloat16* hx, int n, int n_pad, int num_iters, const struct ::himo_deflowpp_weights* w,
              const float4* pt4, float* flow, cudaStream_t stream);
int dec_compact(const float4* pt4, int n, int* pos, int* n_valid, void* scan_scratch, int64_t* valid_idx,
                const float* flow_all, float* flow_valid, cudaStream_t stream);

}  // namespace himo
