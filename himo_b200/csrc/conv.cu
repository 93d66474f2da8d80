// himo_b200/csrc/conv.cu -- H4: dense 2-D convolution of the SeFlow++ backbone as a tcgen05 implicit GEMM.
//
// Replaces the cuDNN/ATen calls behind nn.Conv2d + BatchNorm2d + GELU of `ConvWithNorms`
// (OSF/src/models/basic/__init__.py:76-94) and the bare nn.Conv2d of `UpsampleSkip` / decoder_step4
// (OSF/src/models/basic/unet.py:18-35,130) on the path UNetThreeFrame.forward (unet.py:131-166).
//
// Formulation: D[pixels, Cout] = sum over taps (ky,kx) and input-channel chunks of
//              A_tap[pixels, BK] * W_tap[Cout, BK]^T
//  * activations live in HBM as NHWC 16-bit "planes".  Two planes = split fp16: plane 0 = fp16(x),
//    plane 1 = fp16(x - plane0) (11 + 11 mantissa bits); the kernel then issues three MMAs per k-step
//    (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM; the dropped lo*lo term is ~2^-22 relative), which
//    is what the reference's fp32 path needs for the <=1e-4 flow parity.  Weights are pre-scaled by a
//    power of two (undone exactly in the epilogue, `acc_scale`) so that their low plane stays in
//    fp16's normal range.  One plane = plain bf16 GEMM (the speed mode).
//  * no im2col: for every tap the A tile of 128 output pixels x BK channels is ONE 4-D TMA box
//    (channels, x, y, plane) fetched at (x0*stride + kx - pad, y0*stride + ky - pad); TMA's
//    out-of-bounds zero fill is the convolution padding and its element stride is the conv stride.
//  * tiles land in shared memory in the 64-byte-swizzled K-major layout tcgen05.mma consumes;
//    accumulators stay in tensor memory; one thread issues the MMAs; four warps drain TMEM through
//    tcgen05.ld and fuse bias (folded BN), exact-erf GELU and the hi/lo split into the store.
//  * warp roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 epilogue
//    (TMEM lane quadrant = warp & 3, two warps per quadrant split the columns).
#include <cudaTypedefs.h>

#include "common.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {

constexpr int kConvBM = 128;
constexpr int kConvEpiWarps = 8;
constexpr int kConvThreads = 64 + kConvEpiWarps * 32;

struct ConvParams {
  int tiles_x, tiles_y, n_tiles_n, n_groups;
  int TW, TH;
  int taps, ksize, pad, stride;
  int Cin, cin_off, cin_group_stride, k_chunks;
  const float* bias;
  void* out;
  int out_planes;
  long long out_plane_stride;
  int W_out, Cout_total, cout_off, cout_group_stride;
  int act, out_fp32;
  float acc_scale;
  int flush_iters;
  // GEMM-epilogue extensions used by the FastNSF MLP (csrc/nsf.cu)
  __nv_bfloat16* out_t;            // optional transposed copy: [planes][Cout_total][ld_t], column = pixel
  long long out_t_plane_stride;
  int ld_t;
  const __nv_bfloat16* mask_src;   // optional ReLU-backward mask source, same layout as `out` (2 planes)
  long long mask_plane_stride;
  int mask_planes;
  int b_group_k_stride;            // split-K: K offset of the B operand per group
  long long out_group_pix_stride;  // split-K: output row offset per group
  const int* stop_flag;            // optional device flag: non-zero => the whole launch is a no-op
};

__device__ __forceinline__ float gelu_erf(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

template <int BN, int BK, int P, int STAGES>
struct ConvSmem {
  static constexpr int kABytes = kConvBM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = P * (kABytes + kBBytes);
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + BN * 4 + 1024;  // + barriers + bias + align slack
};

// Accumulation scheme (split mode, P == 2).  tcgen05 adds every MMA into the fp32 TMEM accumulator
// with truncation, so a long accumulation chain drifts by ~0.5 ulp per MMA (measured: error grows
// linearly with K).  To stay fp32-class:
//   * the small cross products hi*lo + lo*hi go to their own accumulator ("cross"), so they never
//     round the large hi*hi sum;
//   * the hi*hi chain is cut every `flush_iters` k-iterations: the MMA warp ping-pongs between two
//     "main" accumulators and the epilogue warps drain the finished one into fp32 registers
//     (round-to-nearest adds) while the tensor core fills the other.
// TMEM columns: main0 [0,BN), main1 [BN,2BN), cross [2BN,3BN).  Single-plane mode uses main0 only.
template <int BN, int BK, int P, int STAGES>
__global__ void __launch_bounds__(kConvThreads)
k_conv_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const ConvParams p) {
  using S = ConvSmem<BN, BK, P, STAGES>;
  constexpr int kUsedCols = P == 2 ? 3 * BN : BN;
  constexpr int kTmemCols = kUsedCols <= 32 ? 32 : kUsedCols <= 64 ? 64 : kUsedCols <= 128 ? 128
                            : kUsedCols <= 256 ? 256 : 512;
  constexpr int kRowBytes = BK * 2;
  constexpr int kHalf = BN / 2;            // columns owned by one epilogue thread
  constexpr int kGroups = kHalf / 16;
  static_assert(kHalf % 16 == 0, "BN must be a multiple of 32");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* acc_empty_bar = acc_full_bar + 2;      // [2]
  uint32_t* tmem_ptr_smem = (uint32_t*)(acc_empty_bar + 2);
  float* bias_s = (float*)(smem + S::kBarOffset + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.stop_flag && *p.stop_flag) return;   // uniform across the grid

  // tile decode: n-tile fastest so CTAs that share an A tile are co-scheduled (L2 reuse)
  int t = blockIdx.x;
  const int nt = t % p.n_tiles_n; t /= p.n_tiles_n;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  const int g = t;
  const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = nt * BN;
  const int k_iters = p.taps * p.k_chunks;
  const int flush = P == 2 ? p.flush_iters : k_iters;
  const int n_chunks = (k_iters + flush - 1) / flush;

  if (warp == 0 && lane == 0) {
    umma::tma_prefetch_desc(&tmA);
    umma::tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(&full_bar[s], 1);
      umma::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(&acc_full_bar[b], 1);
      umma::mbar_init(&acc_empty_bar[b], kConvEpiWarps);
    }
    umma::fence_barrier_init();
  } else if (warp == 1) {
    umma::tmem_alloc(tmem_ptr_smem, kTmemCols);
  } else if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += kConvEpiWarps * 32) bias_s[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      const int cin0 = p.cin_off + g * p.cin_group_stride;
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        umma::mbar_wait(&empty_bar[s], ph ^ 1);
        umma::mbar_arrive_expect_tx(&full_bar[s], S::kStageBytes);
        const int tap = it / p.k_chunks, kc = it - tap * p.k_chunks;
        const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
        uint8_t* a_dst = smem + s * S::kStageBytes;
        uint8_t* b_dst = a_dst + P * S::kABytes;
#pragma unroll
        for (int pl = 0; pl < P; ++pl)
          umma::tma_load_4d(a_dst + pl * S::kABytes, &tmA, &full_bar[s], cin0 + kc * BK,
                            x0 * p.stride + kx - p.pad, y0 * p.stride + ky - p.pad, pl);
#pragma unroll
        for (int pl = 0; pl < P; ++pl)
          umma::tma_load_3d(b_dst + pl * S::kBBytes, &tmB, &full_bar[s],
                            g * p.b_group_k_stride + tap * p.Cin + kc * BK, n0, pl);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      // kind::f16 format codes: 0 = fp16 (split mode, both planes), 1 = bf16 (single-plane mode)
      constexpr uint32_t kFmt = P == 2 ? 0u : 1u;
      constexpr uint32_t idesc = umma::idesc_f16kind_f32(kConvBM, BN, kFmt, kFmt);
      const uint32_t tmem_cross = tmem_base + 2 * BN;
      int it = 0;
      for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int buf = chunk & 1;
        const uint32_t tmem_main = tmem_base + buf * BN;
        if (P == 2) {   // wait until the epilogue has drained this accumulator (2 chunks ago)
          umma::mbar_wait(&acc_empty_bar[buf], ((chunk >> 1) & 1) ^ 1);
          umma::tc_fence_after();
        }
        const int it_begin = it, it_end = min(it + flush, k_iters);
        for (; it < it_end; ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          umma::mbar_wait(&full_bar[s], ph);
          umma::tc_fence_after();
          const uint32_t a_addr = umma::smem_u32(smem + s * S::kStageBytes);
          const uint32_t b_addr = a_addr + P * S::kABytes;
          uint64_t adesc[P], bdesc[P];
#pragma unroll
          for (int pl = 0; pl < P; ++pl) {
            adesc[pl] = umma::smem_desc_kmajor<kRowBytes>(a_addr + pl * S::kABytes);
            bdesc[pl] = umma::smem_desc_kmajor<kRowBytes>(b_addr + pl * S::kBBytes);
          }
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t koff = (uint64_t)(k * 32 >> 4);  // 16 elements = 32 bytes along K
            umma::mma_bf16_ss(tmem_main, adesc[0] + koff, bdesc[0] + koff, idesc,
                              (it != it_begin || k != 0) ? 1u : 0u);
            if (P == 2) {   // hi*lo + lo*hi (lo*lo ~ 2^-22 relative is dropped)
              umma::mma_bf16_ss(tmem_cross, adesc[0] + koff, bdesc[P - 1] + koff, idesc, (it | k) != 0 ? 1u : 0u);
              umma::mma_bf16_ss(tmem_cross, adesc[P - 1] + koff, bdesc[0] + koff, idesc, 1u);
            }
          }
          umma::mma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs have read it
        }
        umma::mma_commit(&acc_full_bar[buf]);  // this chunk's accumulator (and all earlier MMAs) done
      }
    }
  } else {
    // ===================== epilogue (8 warps: 2 per TMEM lane quadrant) =====================
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                 // which half of the BN columns
    const int row = q * 32 + lane;                    // tile row = output pixel within the tile
    const uint32_t lane_col = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * kHalf);
    float acc[kHalf];
#pragma unroll
    for (int j = 0; j < kHalf; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int buf = chunk & 1;
      umma::mbar_wait(&acc_full_bar[buf], (chunk >> 1) & 1);
      umma::tc_fence_after();
#pragma unroll
      for (int gi = 0; gi < kGroups; ++gi) {
        uint32_t r[16];
        umma::tmem_ld_32x16(tmem_base + lane_col + (uint32_t)(buf * BN + gi * 16), r);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[gi * 16 + j] += __uint_as_float(r[j]);
      }
      if (P == 2) {
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&acc_empty_bar[buf]);
      }
    }
    if (P == 2) {   // the last acc_full commit also covers every cross-term MMA
#pragma unroll
      for (int gi = 0; gi < kGroups; ++gi) {
        uint32_t r[16];
        umma::tmem_ld_32x16(tmem_base + lane_col + (uint32_t)(2 * BN + gi * 16), r);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[gi * 16 + j] += __uint_as_float(r[j]);
      }
    }
    const int py = y0 + row / p.TW, px = x0 + row % p.TW;
    const long long pix = (long long)py * p.W_out + px + (long long)g * p.out_group_pix_stride;
    const long long ch0 = (long long)p.cout_off + (long long)g * p.cout_group_stride + n0 + half * kHalf;
#pragma unroll
    for (int gi = 0; gi < kGroups; ++gi) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = __fmaf_rn(acc[gi * 16 + j], p.acc_scale, bias_s[half * kHalf + gi * 16 + j]);
        v[j] = p.act == 0 ? x
               : p.act == 1 ? gelu_erf(x)
               : p.act == 2 ? __fdiv_rn(1.0f, 1.0f + expf(-x))     // torch.sigmoid
               : p.act == 3 ? tanhf(x)                              // torch.tanh
                            : fmaxf(x, 0.f);                        // ReLU
      }
      if (p.mask_src) {   // ReLU backward: pass the gradient where the forward activation was > 0
        const __nv_bfloat16* m = p.mask_src + pix * p.Cout_total + ch0 + gi * 16;
        uint32_t mb[8];
        *(uint4*)&mb[0] = *(const uint4*)m;
        *(uint4*)&mb[4] = *(const uint4*)(m + 8);
        if (p.mask_planes == 2) {
          uint32_t m2[8];
          *(uint4*)&m2[0] = *(const uint4*)(m + p.mask_plane_stride);
          *(uint4*)&m2[4] = *(const uint4*)(m + p.mask_plane_stride + 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) mb[j] |= m2[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if ((mb[j] & 0x00007fffu) == 0u) v[2 * j] = 0.f;
          if ((mb[j] & 0x7fff0000u) == 0u) v[2 * j + 1] = 0.f;
        }
      }
      if (p.out_t) {      // transposed split-plane copy: element (channel, pixel); lanes = consecutive pixels
        const bool split_t = p.out_planes == 2;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          umma::store_split(p.out_t + (ch0 + gi * 16 + j) * (long long)p.ld_t + pix, p.out_t_plane_stride,
                            split_t ? 2 : 1, v[j]);
      }
      if (p.out_fp32) {
        float4* dst = (float4*)((float*)p.out + pix * p.Cout_total + ch0 + gi * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        __nv_bfloat16* o = (__nv_bfloat16*)p.out + pix * p.Cout_total + ch0 + gi * 16;
        const bool split = p.out_planes == 2;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) umma::pack_split2(v[2 * j], v[2 * j + 1], split, hi[j], lo[j]);
        uint4* dst = (uint4*)o;
        dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        if (split) {
          uint4* dst2 = (uint4*)(o + p.out_plane_stride);
          dst2[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          dst2[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ bilinear 2x upsample
// nn.functional.interpolate(scale_factor=2, mode="bilinear", align_corners=False) of
// BilinearDecoder (unet.py:7-16) on an NHWC plane tensor; writes a channel slice of the
// concatenation buffer that u4_u5 reads (torch.cat([u2_res, u3_res], dim=1), unet.py:34).
// src index = (dst + 0.5)/2 - 0.5 clamped at 0 (PyTorch's area_pixel_compute_source_index).
__global__ void __launch_bounds__(256)
k_upsample2x(const __nv_bfloat16* __restrict__ in, int in_planes, long long in_plane_stride, int h, int w,
             int c, __nv_bfloat16* __restrict__ out, int out_planes, long long out_plane_stride,
             int Cout_total, int cout_off) {
  const int c8 = c / 8;
  const long long total = (long long)(2 * h) * (2 * w) * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8) * 8;
    const long long pix = i / c8;
    const int ox = (int)(pix % (2 * w)), oy = (int)(pix / (2 * w));
    const float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
    const float wy[2] = {hy, ly}, wx[2] = {hx, lx};
    float val[4][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const long long off = ((long long)ys[a] * w + xs[b]) * c + cc;
        {
          const uint4 u = *(const uint4*)(in + off);
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = umma::unpack_plane0(uu[k], in_planes == 2);
            val[a * 2 + b][2 * k] = f.x;
            val[a * 2 + b][2 * k + 1] = f.y;
          }
        }
        if (in_planes == 2) {
          const uint4 u = *(const uint4*)(in + in_plane_stride + off);
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&uu[k]));
            val[a * 2 + b][2 * k] += f.x;
            val[a * 2 + b][2 * k + 1] += f.y;
          }
        }
      }
    // same association as ATen's upsample_bilinear2d: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      acc[k] = wy[0] * (wx[0] * val[0][k] + wx[1] * val[1][k]) + wy[1] * (wx[0] * val[2][k] + wx[1] * val[3][k]);
    __nv_bfloat16* o = out + pix * Cout_total + cout_off + cc;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::pack_split2(acc[2 * k], acc[2 * k + 1], out_planes == 2, hi[k], lo[k]);
    *(uint4*)o = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (out_planes == 2) *(uint4*)(o + out_plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled)p;
  }
  return fn;
}

template <int BN, int BK, int P, int STAGES>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, int n_ctas,
                       cudaStream_t stream) {
  using S = ConvSmem<BN, BK, P, STAGES>;
  auto kern = k_conv_umma<BN, BK, P, STAGES>;
  static bool configured = false;
  if (!configured) {
    HIMO_CUDA_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  kern<<<n_ctas, kConvThreads, S::kTotal, stream>>>(tmA, tmB, p);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

}  // namespace himo

using namespace himo;

static int g_flush_iters = 8;
// Tuning knob (process-wide): length of one TMEM accumulation chain in k-iterations of 32 channels.
extern "C" int himo_conv_set_flush_iters(int iters) {
  if (iters < 1) return HIMO_ERR_ARG;
  g_flush_iters = iters;
  return HIMO_OK;
}

extern "C" int himo_conv2d_nhwc(const himo_conv_desc* d, void* stream_) {
  if (!d || !d->in || !d->wgt || !d->out) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  constexpr int BK = 32;
  const int P = d->in_planes;
  if (P != 1 && P != 2) return HIMO_ERR_ARG;
  if (d->ksize != 1 && d->ksize != 3) return HIMO_ERR_UNSUPPORTED;
  if (d->stride != 1 && d->stride != 2) return HIMO_ERR_UNSUPPORTED;
  if (d->Cin % BK || d->Cin_total % 8 || d->Cout_total % 8 || d->cout_off % 8 || d->cin_off % 8)
    return HIMO_ERR_UNSUPPORTED;
  const int pad = d->ksize / 2;
  const int H_out = (d->H_in + 2 * pad - d->ksize) / d->stride + 1;
  const int W_out = (d->W_in + 2 * pad - d->ksize) / d->stride + 1;
  int TW, TH;
  if (W_out >= 128) { TW = 128; TH = 1; }
  else if (W_out == 64) { TW = 64; TH = 2; }
  else if (W_out == 32) { TW = 32; TH = 4; }
  else if (W_out == 16) { TW = 16; TH = 8; }
  else return HIMO_ERR_UNSUPPORTED;
  if (W_out % TW || H_out % TH) return HIMO_ERR_UNSUPPORTED;
  int BN;
  if (d->Cout % 128 == 0) BN = 128;
  else if (d->Cout % 96 == 0) BN = 96;
  else if (d->Cout % 64 == 0) BN = 64;
  else return HIMO_ERR_UNSUPPORTED;
  const int groups = d->n_groups > 0 ? d->n_groups : 1;

  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return HIMO_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin_total, (cuuint64_t)d->W_in, (cuuint64_t)d->H_in, (cuuint64_t)P};
    cuuint64_t strides[3] = {(cuuint64_t)d->Cin_total * 2, (cuuint64_t)d->W_in * d->Cin_total * 2,
                             (cuuint64_t)d->in_plane_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)((TW - 1) * d->stride + 1),
                         (cuuint32_t)((TH - 1) * d->stride + 1), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->in, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return HIMO_ERR_ARG;
  }
  const int taps = d->ksize * d->ksize;
  const long long k_total = d->b_k_total > 0 ? d->b_k_total : (long long)taps * d->Cin;
  {
    cuuint64_t dims[3] = {(cuuint64_t)k_total, (cuuint64_t)d->Cout, (cuuint64_t)P};
    cuuint64_t strides[2] = {(cuuint64_t)k_total * 2, (cuuint64_t)d->Cout * k_total * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d->wgt, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return HIMO_ERR_ARG;
  }
  ConvParams p;
  p.tiles_x = W_out / TW; p.tiles_y = H_out / TH; p.n_tiles_n = d->Cout / BN; p.n_groups = groups;
  p.TW = TW; p.TH = TH;
  p.taps = taps; p.ksize = d->ksize; p.pad = pad; p.stride = d->stride;
  p.Cin = d->Cin; p.cin_off = d->cin_off; p.cin_group_stride = d->cin_group_stride; p.k_chunks = d->Cin / BK;
  p.bias = d->bias; p.out = d->out; p.out_planes = d->out_planes; p.out_plane_stride = d->out_plane_stride;
  p.W_out = W_out; p.Cout_total = d->Cout_total; p.cout_off = d->cout_off;
  p.cout_group_stride = d->cout_group_stride; p.act = d->act; p.out_fp32 = d->out_fp32;
  p.acc_scale = d->acc_scale != 0.f ? d->acc_scale : 1.f;
  p.flush_iters = g_flush_iters;   // k-iterations (2 hi*hi MMAs each) per TMEM accumulation chain
  p.out_t = (__nv_bfloat16*)d->out_t; p.out_t_plane_stride = d->out_t_plane_stride; p.ld_t = d->ld_t;
  p.mask_src = (const __nv_bfloat16*)d->mask_src; p.mask_plane_stride = d->mask_plane_stride;
  p.mask_planes = d->mask_planes;
  p.b_group_k_stride = d->b_group_k_stride; p.out_group_pix_stride = d->out_group_pix_stride;
  p.stop_flag = d->stop_flag;
  const int n_ctas = p.tiles_x * p.tiles_y * p.n_tiles_n * groups;
#define HIMO_CONV_CASE(bn, pp, st) \
  if (BN == bn && P == pp) return launch_conv<bn, BK, pp, st>(tmA, tmB, p, n_ctas, stream);
  HIMO_CONV_CASE(128, 2, 6)
  HIMO_CONV_CASE(96, 2, 7)
  HIMO_CONV_CASE(64, 2, 4)
  HIMO_CONV_CASE(128, 1, 6)
  HIMO_CONV_CASE(96, 1, 6)
  HIMO_CONV_CASE(64, 1, 8)
#undef HIMO_CONV_CASE
  return HIMO_ERR_UNSUPPORTED;
}

extern "C" int himo_upsample2x_nhwc(const void* in, int in_planes, long long in_plane_stride, int h, int w,
                                    int c, void* out, int out_planes, long long out_plane_stride,
                                    int Cout_total, int cout_off, void* stream_) {
  if (!in || !out || c % 8 || Cout_total % 8 || cout_off % 8) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long total = (long long)(2 * h) * (2 * w) * (c / 8);
  const int blocks = (int)((total + 255) / 256 < (long long)kNumSMs * 16 ? (total + 255) / 256 : kNumSMs * 16);
  k_upsample2x<<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)in, in_planes, in_plane_stride, h, w, c,
                                           (__nv_bfloat16*)out, out_planes, out_plane_stride, Cout_total,
                                           cout_off);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
