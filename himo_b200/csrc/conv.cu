// himo_b200/csrc/conv.cu -- H4: dense 2-D convolution of the SeFlow++ backbone as a tcgen05 implicit GEMM.
//
// Replaces the cuDNN/ATen calls behind nn.Conv2d + BatchNorm2d + GELU of `ConvWithNorms`
// (OSF/src/models/basic/__init__.py:76-94) and the bare nn.Conv2d of `UpsampleSkip` / decoder_step4
// (OSF/src/models/basic/unet.py:18-35,130) on the path UNetThreeFrame.forward (unet.py:131-166).
//
// Formulation: D[pixels, Cout] = sum over taps (ky,kx) and input-channel chunks of
//              A_tap[pixels, BK] * W_tap[Cout, BK]^T
//  * activations live in HBM as NHWC bf16 "planes": plane 0 = bf16(x), plane 1 = bf16(x - plane0).
//    With two planes the kernel issues three MMAs per k-step (hi*hi + hi*lo + lo*hi, fp32 accumulate
//    in TMEM): the product error drops to ~2^-17 relative, which is what the reference's fp32 path
//    needs for the <=1e-4 flow parity; with one plane it is a plain bf16 GEMM.
//  * no im2col: for every tap the A tile of 128 output pixels x BK channels is ONE 4-D TMA box
//    (channels, x, y, plane) fetched at (x0*stride + kx - pad, y0*stride + ky - pad); TMA's
//    out-of-bounds zero fill is the convolution padding and its element stride is the conv stride.
//  * tiles land in shared memory in the 64-byte-swizzled K-major layout tcgen05.mma consumes;
//    accumulators stay in tensor memory; one thread issues the MMAs; four warps drain TMEM through
//    tcgen05.ld and fuse bias (folded BN), exact-erf GELU and the hi/lo split into the store.
//  * warp roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue
//    (TMEM lane quadrant = warp & 3).  Two CTAs are co-resident per SM so one tile's epilogue
//    overlaps the other's MMAs.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {

constexpr int kConvBM = 128;
constexpr int kConvThreads = 192;

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

template <int BN, int BK, int P, int STAGES>
__global__ void __launch_bounds__(kConvThreads)
k_conv_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const ConvParams p) {
  using S = ConvSmem<BN, BK, P, STAGES>;
  constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  constexpr int kRowBytes = BK * 2;
  constexpr int kProducts = P == 2 ? 3 : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_full_bar + 1);
  float* bias_s = (float*)(smem + S::kBarOffset + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile decode: n-tile fastest so CTAs that share an A tile are co-scheduled (L2 reuse)
  int t = blockIdx.x;
  const int nt = t % p.n_tiles_n; t /= p.n_tiles_n;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  const int g = t;
  const int x0 = tx * p.TW, y0 = ty * p.TH, n0 = nt * BN;
  const int k_iters = p.taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    umma::tma_prefetch_desc(&tmA);
    umma::tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(&full_bar[s], 1);
      umma::mbar_init(&empty_bar[s], 1);
    }
    umma::mbar_init(tmem_full_bar, 1);
    umma::fence_barrier_init();
  } else if (warp == 1) {
    umma::tmem_alloc(tmem_ptr_smem, kTmemCols);
  } else if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) bias_s[i] = p.bias ? __ldg(p.bias + n0 + i) : 0.f;
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
          umma::tma_load_3d(b_dst + pl * S::kBBytes, &tmB, &full_bar[s], tap * p.Cin + kc * BK, n0, pl);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma::idesc_bf16_f32(kConvBM, BN);
      for (int it = 0; it < k_iters; ++it) {
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
          const uint64_t koff = (uint64_t)(k * 32 >> 4);  // 16 bf16 = 32 bytes along K
#pragma unroll
          for (int pr = 0; pr < kProducts; ++pr) {
            // products: hi*hi, hi*lo, lo*hi (lo*lo ~ 2^-18 relative is dropped)
            const int pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;
            umma::mma_bf16_ss(tmem_base, adesc[pa] + koff, bdesc[pb] + koff, idesc,
                              (it | k | pr) != 0 ? 1u : 0u);
          }
        }
        umma::mma_commit(&empty_bar[s]);   // frees the smem stage once these MMAs have read it
      }
      umma::mma_commit(tmem_full_bar);     // accumulator complete -> epilogue
    }
  } else {
    // ===================== epilogue (4 warps) =====================
    umma::mbar_wait(tmem_full_bar, 0);
    umma::tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;          // tile row = output pixel within the tile
    const int py = y0 + row / p.TW, px = x0 + row % p.TW;
    const long long pix = (long long)py * p.W_out + px;
    const long long ch0 = (long long)p.cout_off + (long long)g * p.cout_group_stride + n0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      umma::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      umma::tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(r[j]) + bias_s[c * 32 + j];
        v[j] = p.act == 1 ? gelu_erf(x) : x;
      }
      if (p.out_fp32) {
        float4* dst = (float4*)((float*)p.out + pix * p.Cout_total + ch0 + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      } else {
        __nv_bfloat16* o = (__nv_bfloat16*)p.out + pix * p.Cout_total + ch0 + c * 32;
        uint32_t hi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hi[j] = umma::pack_bf16x2(v[2 * j], v[2 * j + 1]);
        uint4* dst = (uint4*)o;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        if (p.out_planes == 2) {
          uint32_t lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float h0 = __uint_as_float(hi[j] << 16), h1 = __uint_as_float(hi[j] & 0xffff0000u);
            lo[j] = umma::pack_bf16x2(v[2 * j] - h0, v[2 * j + 1] - h1);
          }
          uint4* dst2 = (uint4*)(o + p.out_plane_stride);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst2[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
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
#pragma unroll
        for (int k = 0; k < 8; ++k) val[a * 2 + b][k] = 0.f;
        for (int pl = 0; pl < in_planes; ++pl) {
          const uint4 u = *(const uint4*)(in + pl * in_plane_stride + off);
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            val[a * 2 + b][2 * k] += __uint_as_float(uu[k] << 16);
            val[a * 2 + b][2 * k + 1] += __uint_as_float(uu[k] & 0xffff0000u);
          }
        }
      }
    // same association as ATen's upsample_bilinear2d: h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      acc[k] = wy[0] * (wx[0] * val[0][k] + wx[1] * val[1][k]) + wy[1] * (wx[0] * val[2][k] + wx[1] * val[3][k]);
    __nv_bfloat16* o = out + pix * Cout_total + cout_off + cc;
    uint32_t hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) hi[k] = umma::pack_bf16x2(acc[2 * k], acc[2 * k + 1]);
    *(uint4*)o = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (out_planes == 2) {
      uint32_t lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        lo[k] = umma::pack_bf16x2(acc[2 * k] - __uint_as_float(hi[k] << 16),
                                  acc[2 * k + 1] - __uint_as_float(hi[k] & 0xffff0000u));
      *(uint4*)(o + out_plane_stride) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
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
  const long long k_total = (long long)taps * d->Cin;
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
  const int n_ctas = p.tiles_x * p.tiles_y * p.n_tiles_n * groups;
#define HIMO_CONV_CASE(bn, pp, st) \
  if (BN == bn && P == pp) return launch_conv<bn, BK, pp, st>(tmA, tmB, p, n_ctas, stream);
  HIMO_CONV_CASE(128, 2, 3)
  HIMO_CONV_CASE(96, 2, 3)
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
