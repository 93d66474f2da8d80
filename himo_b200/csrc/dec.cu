// himo_b200/csrc/dec.cu -- H4 back end: the per-point ConvGRU flow decoder of SeFlow++.
//
// Replaces ConvGRUDecoder.forward_single + ConvGRU.forward (OSF/src/models/basic/decoder.py:177-237).
// The three Conv1d(288->192, k=1) of the GRU and Linear(288->48) are GEMMs over the points; they run
// on the same tcgen05 kernel as the backbone (csrc/conv.cu, the points viewed as a [N/128,128] image,
// 1x1 taps) with sigmoid / tanh / GELU fused into the TMEM epilogue.  This file holds the memory-bound
// glue around them: the gather of the 2x96 pillar vectors + offset encoder, r*h, the state update and
// the final Linear(48->3).  Points are NOT compacted: dropped points (key < 0) ride along as zero rows
// so that no size ever has to come back to the host.
#include "common.cuh"
#include "dec.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {


using umma::store_split;

// 8 consecutive channels per thread: 2 x float4 in, one 16-byte store per plane out
__device__ __forceinline__ void store_split8(__nv_bfloat16* dst, long long plane_stride, int planes, const float* v) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) umma::pack_split2(v[2 * k], v[2 * k + 1], planes == 2, hi[k], lo[k]);
  *(uint4*)dst = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  if (planes == 2) *(uint4*)(dst + plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// one thread per (point, 8-channel chunk) of the 288 = [before(96) | after(96) | offset feature(96)] channels: 36 chunks per
// point, consecutive threads = consecutive chunks, so every global access is a 16/32-byte vector and all lanes work (with one
// warp per point the second trip over the 36 chunks kept 4 of 32 lanes busy).
__global__ void __launch_bounds__(256)
k_dec_gather(DecGatherArgs a) {
  const long long total = (long long)a.n_pad * 36;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / 36);
    const int ch = (int)(t - (long long)i * 36);
    float4 p = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (i < a.n) p = a.pt4[i];
    const int key = __float_as_int(p.w);
    float* h = a.h32 ? a.h32 + (size_t)i * 192 : nullptr;
    __nv_bfloat16* hx = a.hx_planes + (size_t)i * 288;
    __nv_bfloat16* rhx = a.rhx_planes ? a.rhx_planes + (size_t)i * 288 : nullptr;
    const int c0 = ch * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (key >= 0) {
      if (ch < 12) {
        // before_pseudoimage[:, y, x]: the exact fp32 voxel feature of frame f, zero where that frame is empty
        const int f = ch >> 2;
        const unsigned* bm = a.bitmap + (size_t)f * a.n_words;
        if ((__ldg(bm + (key >> 5)) >> (key & 31)) & 1u) {
          const int r = bitmap_rank_lb(bm, a.word_prefix + (size_t)f * a.n_words, key);
          const float4* src = (const float4*)(a.voxel_feats + ((size_t)f * a.n_max + r) * 32 + (ch & 3) * 8);
          *(float4*)&v[0] = __ldg(src); *(float4*)&v[4] = __ldg(src + 1);
        }
      } else if (ch < 24) {
        const float4* src = (const float4*)(a.after + (size_t)key * a.c_after + (c0 - 96));   // after[:, y, x]
        *(float4*)&v[0] = __ldg(src); *(float4*)&v[4] = __ldg(src + 1);
      } else {
        // point_offsets = p - ((c * voxel_size + min) + voxel_size/2), every step rounded to fp32
        // (DynamicVoxelizer._get_point_offsets, encoder.py:506-523)
        const int cy = key / a.gx, cx = key - cy * a.gx;
        const float ox = p.x - __fadd_rn(__fadd_rn(__fmul_rn((float)cx, a.vx), a.x_min), a.hx);
        const float oy = p.y - __fadd_rn(__fadd_rn(__fmul_rn((float)cy, a.vy), a.y_min), a.hy);
        const float oz = p.z - __fadd_rn(__fadd_rn(__fmul_rn(0.f, a.vz), a.z_min), a.hz);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = c0 - 192 + k;
          float tt = __ldg(a.w_off + c * 3) * ox;
          tt = fmaf(__ldg(a.w_off + c * 3 + 1), oy, tt);
          tt = fmaf(__ldg(a.w_off + c * 3 + 2), oz, tt);
          v[k] = tt + __ldg(a.b_off + c);
        }
      }
    }
    if (h && ch < 24) { *(float4*)(h + c0) = *(float4*)&v[0]; *(float4*)(h + c0 + 4) = *(float4*)&v[4]; }
    store_split8(hx + c0, a.plane_stride, a.planes, v);
    if (rhx && ch >= 24) store_split8(rhx + c0, a.plane_stride, a.planes, v);
  }
}

// rhx[:, :192] = split(r * h)      (ConvGRU.forward: rh_x = cat([r*h, x]), decoder.py:189)
__global__ void __launch_bounds__(256)
k_dec_rh(const float* __restrict__ zr, const float* __restrict__ h32, int n_pad,
         __nv_bfloat16* __restrict__ rhx, int planes, long long plane_stride) {
  const long long total = (long long)n_pad * 24;          // 192 / 8 chunks per point
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / 24;
    const int c = (int)(t - i * 24) * 8;
    const float4 r0 = *(const float4*)(zr + i * 384 + 192 + c), r1 = *(const float4*)(zr + i * 384 + 196 + c);
    const float4 h0 = *(const float4*)(h32 + i * 192 + c), h1 = *(const float4*)(h32 + i * 192 + c + 4);
    const float v[8] = {r0.x * h0.x, r0.y * h0.y, r0.z * h0.z, r0.w * h0.w, r1.x * h1.x, r1.y * h1.y, r1.z * h1.z, r1.w * h1.w};
    store_split8(rhx + i * 288 + c, plane_stride, planes, v);
  }
}

// h = (1 - z) * h + z * q          (decoder.py:192)
__global__ void __launch_bounds__(256)
k_dec_update(const float* __restrict__ zr, const float* __restrict__ q, float* __restrict__ h32, int n_pad,
             __nv_bfloat16* __restrict__ hx, int planes, long long plane_stride) {
  const long long total = (long long)n_pad * 24;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / 24;
    const int c = (int)(t - i * 24) * 8;
    float z[8], qq[8], h[8], hn[8];
    *(float4*)&z[0] = *(const float4*)(zr + i * 384 + c); *(float4*)&z[4] = *(const float4*)(zr + i * 384 + c + 4);
    *(float4*)&qq[0] = *(const float4*)(q + i * 192 + c); *(float4*)&qq[4] = *(const float4*)(q + i * 192 + c + 4);
    *(float4*)&h[0] = *(const float4*)(h32 + i * 192 + c); *(float4*)&h[4] = *(const float4*)(h32 + i * 192 + c + 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) hn[k] = __fadd_rn(__fmul_rn(1.0f - z[k], h[k]), __fmul_rn(z[k], qq[k]));
    *(float4*)(h32 + i * 192 + c) = *(float4*)&hn[0];
    *(float4*)(h32 + i * 192 + c + 4) = *(float4*)&hn[4];
    store_split8(hx + i * 288 + c, plane_stride, planes, hn);
  }
}

// flow = Linear(48->3)(y); y already holds GELU(Linear(288->48)) from the GEMM epilogue.
__global__ void __launch_bounds__(256)
k_dec_out(const float* __restrict__ y, int y_stride, const float4* __restrict__ pt4, int n,
          const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ flow) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float o[3] = {0.f, 0.f, 0.f};
    if (__float_as_int(pt4[i].w) >= 0) {
      const float4* yr = (const float4*)(y + (size_t)i * y_stride);
      float yv[48];
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        float4 t = yr[k];
        yv[4 * k] = t.x; yv[4 * k + 1] = t.y; yv[4 * k + 2] = t.z; yv[4 * k + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float acc = __ldg(w2 + j * 48) * yv[0];
#pragma unroll
        for (int k = 1; k < 48; ++k) acc = fmaf(__ldg(w2 + j * 48 + k), yv[k], acc);
        o[j] = acc + __ldg(b2 + j);
      }
    }
    flow[3 * (size_t)i] = o[0]; flow[3 * (size_t)i + 1] = o[1]; flow[3 * (size_t)i + 2] = o[2];
  }
}

// pose_flow + flow assembly and ordered compaction helpers -------------------------------------
__global__ void __launch_bounds__(256)
k_valid_flags_write(const float4* __restrict__ pt4, int n, const int* __restrict__ pos,
                    int64_t* __restrict__ valid_idx, const float* __restrict__ flow_all,
                    float* __restrict__ flow_valid) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (__float_as_int(pt4[i].w) >= 0) {
      const int j = pos[i];
      valid_idx[j] = i;
      flow_valid[3 * (size_t)j] = flow_all[3 * (size_t)i];
      flow_valid[3 * (size_t)j + 1] = flow_all[3 * (size_t)i + 1];
      flow_valid[3 * (size_t)j + 2] = flow_all[3 * (size_t)i + 2];
    }
  }
}

struct MapValidPt4 {
  const float4* p;
  __device__ int operator()(int i) const { return __float_as_int(p[i].w) >= 0 ? 1 : 0; }
};

}  // namespace himo

using namespace himo;

// C++-level entry points used by deflowpp.cu (same shared object).
namespace himo {

int dec_gather(const DecGatherArgs& a, cudaStream_t stream) {
  k_dec_gather<<<(int)min(ceil_div_ll((long long)a.n_pad * 36, 256), (long long)kNumSMs * 16), 256, 0, stream>>>(a);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
int dec_rh(const float* zr, const float* h32, int n_pad, __nv_bfloat16* rhx, int planes, long long ps,
           cudaStream_t stream) {
  k_dec_rh<<<kNumSMs * 8, 256, 0, stream>>>(zr, h32, n_pad, rhx, planes, ps);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
int dec_update(const float* zr, const float* q, float* h32, int n_pad, __nv_bfloat16* hx, int planes,
               long long ps, cudaStream_t stream) {
  k_dec_update<<<kNumSMs * 8, 256, 0, stream>>>(zr, q, h32, n_pad, hx, planes, ps);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
int dec_out(const float* y, int y_stride, const float4* pt4, int n, const float* w2, const float* b2,
            float* flow, cudaStream_t stream) {
  if (n <= 0) return HIMO_OK;
  k_dec_out<<<min(ceil_div(n, 256), kNumSMs * 8), 256, 0, stream>>>(y, y_stride, pt4, n, w2, b2, flow);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
// ordered compaction of the valid points: valid_idx [n_valid] i64, flow_valid [n_valid,3], n_valid (device)
int dec_compact(const float4* pt4, int n, int* pos, int* n_valid, void* scan_scratch, int64_t* valid_idx,
                const float* flow_all, float* flow_valid, cudaStream_t stream) {
  HIMO_CUDA_RET(scan_exclusive(MapValidPt4{pt4}, pos, n, nullptr, n_valid, scan_scratch, stream));
  if (n <= 0) return HIMO_OK;
  k_valid_flags_write<<<min(ceil_div(n, 256), kNumSMs * 8), 256, 0, stream>>>(pt4, n, pos, valid_idx, flow_all,
                                                                             flow_valid);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

}  // namespace himo
