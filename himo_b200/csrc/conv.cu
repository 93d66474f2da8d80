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
constexpr int kConvBK = 32;               // input channels per pipeline stage (64-byte swizzled rows)
constexpr int kConvRowB = 64;
constexpr int kConvEpiWarps = 8;
constexpr int kConvThreads = 64 + kConvEpiWarps * 32;

struct ConvParams {
  int tiles_x, tiles_y, n_tiles_n, n_groups, total_tiles;
  int TW, TH;
  int taps, ksize, pad, stride;
  int Cin, cin_off, cin_group_stride, k_chunks;
  const float* bias;
  void* out;
  int out_planes;
  long long out_plane_stride;
  int W_out, Cout_total, cout_off, cout_group_stride;
  int Cout;                         // output channels of one group (bias length)
  int act, out_fp32;
  float acc_scale;
  int flush_stages;
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
  // fused ConvGRU epilogues (act 5 / 6), see the epilogue
  float* aux_h;                    // [pixels][aux_ld] fp32 hidden state h
  float* aux_z;                    // [pixels][aux_ld] fp32 update gate z
  int aux_ld;
  __nv_bfloat16* out2;             // split-plane operand buffer of the NEXT GEMM ([planes][pixels][out2_ld])
  long long out2_plane_stride;
  int out2_ld;
  // composed 1x1 -> 3x3 convolutions (deflowpp.cu): bias per border class [3][3][Cout] (row class, column class;
  // 0 = first row/column, 1 = interior, 2 = last), used instead of `bias` for pixels on the image border
  const float* border_bias;
  int H_out;
  long long* dbg;                  // optional per-tile clock64() trace [CTA][32 tiles][8] (himo_conv_set_debug_buffer)
};

// sigmoid / tanh through ex2.approx + fast division: abs error ~2e-7, far below the 1e-4 flow budget, and
// 4x fewer instructions than expf + IEEE divide (the GRU GEMMs were epilogue-bound on them)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
// Exact-erf GELU (nn.GELU(), basic/__init__.py:76-94) without erff's branches: 0.5 x (1 + erf(x/sqrt2)) = 0.5 x erfc(-x/sqrt2),
// erfc(t) = 2^p(t) for t = |x|/sqrt2 in [0, 5] with p a degree-8 minimax fit of log2(erfc) weighted for uniform ABSOLUTE
// error of erfc (5.9e-8 in fp32 Horner form + ex2.approx's 2^-22 relative).  |GELU error| <= 3e-7 max(|x|, 1) -- below
// the 1.1e-6 of torch's own fp32 CPU GELU against the exact function on [-12, 12] -- in 13 instructions instead of ~35
// (the epilogue warps were issue-bound on erff: profiles/r02_conv_tile_trace.txt).
__device__ __forceinline__ float gelu_erf(float v) {
  const float t = fminf(fabsf(v) * 0.70710678118654752440f, 5.0f);
  float p = -4.535873086e-05f;
  p = __fmaf_rn(p, t, 4.455077578e-04f);
  p = __fmaf_rn(p, t, -1.489439164e-03f);
  p = __fmaf_rn(p, t, -7.746376796e-04f);
  p = __fmaf_rn(p, t, 2.825369127e-02f);
  p = __fmaf_rn(p, t, -1.484816223e-01f);
  p = __fmaf_rn(p, t, -9.184163809e-01f);
  p = __fmaf_rn(p, t, -1.627908587e+00f);
  float e;                                            // 2^(p t) in [2^-46, 1]: one MUFU.EX2, no range fix-up needed
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p * t));
  const float h = 0.5f * v * e;                       // 0.5 x erfc(|x|/sqrt2)
  return v > 0.f ? v - h : h;
}

// Pipeline stage = one A load shared by NX taps + the NX weight tiles.
//   NX == 3 ("halo" mode, 3x3 stride-1 convs on rows >= 128 px): the A tile is ONE haloed row of
//     128+2 pixels x 32 channels; the taps kx = 0,1,2 read it through three descriptors whose start
//     address is shifted by one 64-byte row each (the tensor core swizzles on absolute shared-memory
//     address bits, like TMA -- profiles/r01_exp_shifted_descriptor.txt), so the activation tile is
//     fetched from L2 three times per output tile instead of nine.
//   NX == 1: one tap per stage (1x1 convs, stride-2 convs, 64-px rows, plain GEMMs).
// CG == 2: a pair of CTAs (thread-block cluster of 2) computes a 256-pixel x BN tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 pixels of A and only HALF of the weight rows, the
// leader CTA issues the MMAs for both, and every weight byte is fetched from L2 once per 256 pixels.
// WR > 0 ("weights resident"): the layer's whole weight tensor -- WR = taps * k_chunks tiles of BN rows x 32
// channels per plane -- is loaded into shared memory ONCE per CTA and every pipeline stage carries only the
// activation tile.  For the 64-channel encoder layers the weights (74 / 147 KB) were re-fetched from L2 for every
// 128-pixel tile and out-weighed the activation traffic 3:2.
// TS ("TMA store"): the epilogue streams TMEM -> bias / activation -> a 128-byte-swizzled staging tile in shared memory
// (64 channels x 128 pixels x both planes = 32 KB) and ONE thread hands it to TMA, instead of every thread storing
// 16-byte pieces 2*Cout_total bytes apart (32 half-filled sectors per warp store: the tile trace showed ~10 k cycles of
// stores per 128x128 tile, profiles/r02_conv_tile_trace_enc2_after.txt).  Requires a single accumulation chunk per tile.
template <int BN, int P, int NX, int CG, int WR = 0, int TS = 0>
struct ConvCfg {
  static constexpr int kARows = NX == 3 ? 136 : 128;                 // smem rows reserved per A plane
  static constexpr int kARowsTx = NX == 3 ? 130 : 128;               // rows TMA really writes
  static constexpr int kABytes = kARows * kConvRowB;
  static constexpr int kBRows = BN / CG;                             // weight rows staged by this CTA
  static constexpr int kBBytes = kBRows * kConvRowB;
  static constexpr int kBResBytes = WR * P * kBBytes;
  static constexpr int kStageBytes = WR ? P * kABytes : P * kABytes + NX * P * kBBytes;
  static constexpr int kTxBytes = WR ? P * kARowsTx * kConvRowB : P * kARowsTx * kConvRowB + NX * P * kBBytes;
  static constexpr int kOutStageBytes = TS ? 2 * 2 * 128 * 128 : 0; // two [plane][128 px][64 ch] staging tiles of the TMA-store epilogue
  static constexpr int kStagesRaw = (219 * 1024 - kBResBytes - kOutStageBytes) / kStageBytes;
  static constexpr int STAGES = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kBResOffset = STAGES * kStageBytes;          // (TMA and tcgen05 both swizzle on absolute address bits)
  static constexpr int kOutOffset = (kBResOffset + kBResBytes + 1023) / 1024 * 1024;   // 1024-aligned (SWIZZLE_128B) from a 1024-aligned base
  static constexpr int kBarOffset = TS ? kOutOffset + kOutStageBytes : kBResOffset + kBResBytes;
  static_assert(WR == 0 || CG == 1, "resident weights are implemented for single-CTA tiles");
  static constexpr int kBiasOffset = kBarOffset + 512;               // fp32 bias vector staged once per CTA
  static constexpr int kMaxBias = 1024;
  static constexpr int kAlign = TS ? 1024 : 128;
  static constexpr int kTotal = kBiasOffset + kMaxBias * 4 + kAlign;   // barriers + bias + alignment slack
  static_assert(!TS || (P == 2 && BN % 64 == 0 && WR == 0), "TMA-store epilogue: split planes, 64-channel chunks");
  // split mode: 2 main + 1 or 2 cross accumulators, then 8 staging slots x 16 columns for the A slices in TMEM
  static constexpr int kCrossBufs = (P == 2 && 4 * BN <= 512) ? 2 : 1;
  static constexpr int kAccCols = P == 2 ? (2 + kCrossBufs) * BN : 2 * BN;
  static constexpr int kUsedCols = kAccCols;
  static constexpr int kTmemCols = kUsedCols <= 64 ? 64 : kUsedCols <= 128 ? 128 : kUsedCols <= 256 ? 256 : 512;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
};

// Tile epilogue for one thread: kHalfT accumulator columns of one output pixel.
template <int ACT, int kHalfT>
__device__ __forceinline__ void conv_epilogue(const float* acc, const ConvParams& p, long long pix, long long ch0,
                                              const float* bias) {
#pragma unroll
for (int gi = 0; gi < kHalfT / 16; ++gi) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x = __fmaf_rn(acc[gi * 16 + j], p.acc_scale, bias ? __ldg(bias + gi * 16 + j) : 0.f);
    if (ACT == 1) v[j] = gelu_erf(x);
    else if (ACT == 2) v[j] = fast_sigmoid(x);                       // torch.sigmoid
    else if (ACT == 3) v[j] = fast_tanh(x);                          // torch.tanh
    else if (ACT == 4) v[j] = fmaxf(x, 0.f);                         // ReLU
    else v[j] = x;                                                   // 0: none; 5/6: fused ConvGRU below
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
  if (ACT >= 5) {
    // ConvGRU fused epilogues (OSF/src/models/basic/decoder.py:185-192), column c of this GEMM:
    //  act 5 (zr GEMM, 2*aux_ld columns): c <  aux_ld: z = sigmoid(.) -> aux_z
    //                                     c >= aux_ld: r = sigmoid(.), out2[:, c-aux_ld] = split(r*h)
    //  act 6 (q GEMM): q = tanh(.), h <- (1-z)*h + z*q -> aux_h and out2[:, c] = split(h)
    const long long cbase = ch0 + gi * 16;
    const bool split2 = p.out_planes == 2;
    if (ACT == 5) {
      if (cbase < p.aux_ld) {
        float4* zp = (float4*)(p.aux_z + pix * p.aux_ld + cbase);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 t;
          t.x = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j])); t.y = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j + 1]));
          t.z = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j + 2])); t.w = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j + 3]));
          zp[j] = t;
        }
      } else {
        const long long c2 = cbase - p.aux_ld;
        const float4* hp = (const float4*)(p.aux_h + pix * p.aux_ld + c2);
        float rh[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 h4 = hp[j];
          rh[4 * j] = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j])) * h4.x;
          rh[4 * j + 1] = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j + 1])) * h4.y;
          rh[4 * j + 2] = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j + 2])) * h4.z;
          rh[4 * j + 3] = __fdiv_rn(1.0f, 1.0f + expf(-v[4 * j + 3])) * h4.w;
        }
        __nv_bfloat16* o2 = p.out2 + pix * p.out2_ld + c2;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) umma::pack_split2(rh[2 * j], rh[2 * j + 1], split2, hi[j], lo[j]);
        ((uint4*)o2)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        ((uint4*)o2)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        if (split2) {
          ((uint4*)(o2 + p.out2_plane_stride))[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          ((uint4*)(o2 + p.out2_plane_stride))[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    } else {
      float4* hp = (float4*)(p.aux_h + pix * p.aux_ld + cbase);
      const float4* zp = (const float4*)(p.aux_z + pix * p.aux_ld + cbase);
      float hn[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 h4 = hp[j], z4 = zp[j];
        hn[4 * j] = __fadd_rn(__fmul_rn(1.0f - z4.x, h4.x), __fmul_rn(z4.x, tanhf(v[4 * j])));
        hn[4 * j + 1] = __fadd_rn(__fmul_rn(1.0f - z4.y, h4.y), __fmul_rn(z4.y, tanhf(v[4 * j + 1])));
        hn[4 * j + 2] = __fadd_rn(__fmul_rn(1.0f - z4.z, h4.z), __fmul_rn(z4.z, tanhf(v[4 * j + 2])));
        hn[4 * j + 3] = __fadd_rn(__fmul_rn(1.0f - z4.w, h4.w), __fmul_rn(z4.w, tanhf(v[4 * j + 3])));
        hp[j] = make_float4(hn[4 * j], hn[4 * j + 1], hn[4 * j + 2], hn[4 * j + 3]);
      }
      __nv_bfloat16* o2 = p.out2 + pix * p.out2_ld + cbase;
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) umma::pack_split2(hn[2 * j], hn[2 * j + 1], split2, hi[j], lo[j]);
      ((uint4*)o2)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      ((uint4*)o2)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      if (split2) {
        ((uint4*)(o2 + p.out2_plane_stride))[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        ((uint4*)(o2 + p.out2_plane_stride))[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    }
    continue;
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

// Common-path tile epilogue for 16 accumulator columns of one output pixel: scale + bias, activation, then
// either 16 fp32 values or the split 16-bit planes.  Deliberately NOT inlined: the fully unrolled epilogue of
// a 128-column tile was ~50 KB of straight-line code per activation, executed once per tile, and the
// profile (profiles/r01_enc2_source_stalls.txt) showed the epilogue warps starving on instruction fetch
// (stall_no_inst).  One ~700-instruction body per activation stays resident in the instruction caches.
struct EpiOut {
  void* out;                 // address of this pixel's first channel of the group (plane 0)
  long long plane_stride_b;  // bytes between the hi and lo planes (split output)
  float scale;
  int mode;                  // 0 = fp32, 1 = one bf16 plane, 2 = split fp16 planes
  // optional (FastNSF GEMMs): ReLU-backward mask source (layout of `out`) and transposed split-plane copy
  const __nv_bfloat16* mask;       // this pixel's first channel of the group, or nullptr
  long long mask_plane_stride; int mask_planes;
  __nv_bfloat16* out_t;            // element (first channel of the group, this pixel), or nullptr
  long long out_t_plane_stride; long long ld_t; int t_planes;
};
template <int ACT>
__device__ __noinline__ void epi_store16(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7,
                                         float a8, float a9, float a10, float a11, float a12, float a13, float a14,
                                         float a15, const float* __restrict__ bias_s, EpiOut o) {
  float v[16] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15};
  float b[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) *(float4*)&b[4 * j] = *(const float4*)(bias_s + 4 * j);   // shared-memory broadcast
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float x = __fmaf_rn(v[j], o.scale, b[j]);
    if (ACT == 1) v[j] = gelu_erf(x);
    else if (ACT == 2) v[j] = fast_sigmoid(x);
    else if (ACT == 3) v[j] = fast_tanh(x);
    else if (ACT == 4) v[j] = fmaxf(x, 0.f);
    else v[j] = x;
  }
  if (o.mask) {   // ReLU backward: pass the gradient where the forward activation was > 0
    uint32_t mb[8];
    *(uint4*)&mb[0] = *(const uint4*)o.mask;
    *(uint4*)&mb[4] = *(const uint4*)(o.mask + 8);
    if (o.mask_planes == 2) {
      uint32_t m2[8];
      *(uint4*)&m2[0] = *(const uint4*)(o.mask + o.mask_plane_stride);
      *(uint4*)&m2[4] = *(const uint4*)(o.mask + o.mask_plane_stride + 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) mb[j] |= m2[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((mb[j] & 0x00007fffu) == 0u) v[2 * j] = 0.f;
      if ((mb[j] & 0x7fff0000u) == 0u) v[2 * j + 1] = 0.f;
    }
  }
  if (o.out_t) {   // transposed copy: element (channel, pixel); the lanes of a warp hold consecutive pixels
#pragma unroll
    for (int j = 0; j < 16; ++j)
      umma::store_split(o.out_t + j * o.ld_t, o.out_t_plane_stride, o.t_planes, v[j]);
  }
  if (o.mode == 0) {
    float4* dst = (float4*)o.out;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    const bool split = o.mode == 2;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) umma::pack_split2(v[2 * j], v[2 * j + 1], split, hi[j], lo[j]);
    uint4* dst = (uint4*)o.out;
    dst[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dst[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    if (split) {
      uint4* dst2 = (uint4*)((char*)o.out + o.plane_stride_b);
      dst2[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      dst2[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
  }
}

template <int ACT, int kHalfT>
__device__ __forceinline__ void conv_epilogue_fast(const float* acc, const ConvParams& p, long long pix, long long ch0,
                                                   const float* bias_s) {
  EpiOut o;
  o.scale = p.acc_scale;
  o.mode = p.out_fp32 ? 0 : (p.out_planes == 2 ? 2 : 1);
  o.plane_stride_b = p.out_plane_stride * 2;
  o.mask_plane_stride = p.mask_plane_stride; o.mask_planes = p.mask_planes;
  o.out_t_plane_stride = p.out_t_plane_stride; o.ld_t = p.ld_t; o.t_planes = p.out_planes == 2 ? 2 : 1;
  const long long elem = pix * p.Cout_total + ch0;
#pragma unroll
  for (int gi = 0; gi < kHalfT / 16; ++gi) {
    o.out = p.out_fp32 ? (void*)((float*)p.out + elem + gi * 16) : (void*)((__nv_bfloat16*)p.out + elem + gi * 16);
    o.mask = p.mask_src ? p.mask_src + elem + gi * 16 : nullptr;
    o.out_t = p.out_t ? p.out_t + (ch0 + gi * 16) * (long long)p.ld_t + pix : nullptr;
    const float* a = acc + gi * 16;
    epi_store16<ACT>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14],
                     a[15], bias_s + gi * 16, o);
  }
}

// Accumulation scheme.  tcgen05 adds every MMA into the fp32 TMEM accumulator with truncation, so a long
// accumulation chain drifts by ~0.5 ulp per MMA (measured: error grows linearly with K).  To stay fp32-class
//   * split mode: the small cross products hi*lo + lo*hi go to their own accumulator ("cross") and never
//     round the large hi*hi sum;
//   * the hi*hi chain is cut every `flush_stages` stages: the MMA warp ping-pongs between two "main"
//     accumulators and the epilogue warps drain the finished one into fp32 registers (round-to-nearest)
//     while the tensor core fills the other.
// TMEM columns: main0 [0,BN), main1 [BN,2BN), cross0 [2BN,3BN), cross1 [3BN,4BN): the cross accumulator is
// double-buffered by tile parity so that the MMAs of tile i+1 never wait for the drain of tile i (with short-K
// tiles -- 6 stages for the 64-channel layers, 4 for the FastNSF GEMMs -- that bubble was ~20 % of a tile).
// The kernel is persistent: one CTA per SM walks the tile list, barrier phases run across tiles, and the
// store epilogue of tile i overlaps the main loop of tile i+1.
template <int BN, int P, int NX, int CG, int WR = 0, int TS = 0>
__global__ void __launch_bounds__(kConvThreads, 1)
k_conv_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmO, const ConvParams p) {
  using C = ConvCfg<BN, P, NX, CG, WR, TS>;
  constexpr int STAGES = C::STAGES;
  constexpr int kHalf = BN / 2;            // columns owned by one epilogue thread
  constexpr int kGroups = kHalf / 16;
  static_assert(kHalf % 16 == 0, "BN must be a multiple of 32");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + (C::kAlign - 1)) & ~(uintptr_t)(C::kAlign - 1));
  uint64_t* full_bar = (uint64_t*)(smem + C::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full_bar = empty_bar + STAGES;     // [2]
  uint64_t* acc_empty_bar = acc_full_bar + 2;      // [2]
  uint64_t* cross_empty_bar = acc_empty_bar + 2;   // [2]
  uint64_t* bres_bar = cross_empty_bar + 2;       // [1] resident weights landed (WR mode)
  uint32_t* tmem_ptr_smem = (uint32_t*)(bres_bar + 1);
  float* bias_smem = (float*)(smem + C::kBiasOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.stop_flag) {                         // the flag is written by an earlier kernel of the stream
    pdl_wait();
    if (*p.stop_flag) return;                // uniform across the grid
  }
  const uint32_t cta_rank = CG == 2 ? umma::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int n_workers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;    // CTAs (or CTA pairs)
  const int worker = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;

  const int k_iters = (p.taps / NX) * p.k_chunks;          // pipeline stages per tile
  const int flush = P == 2 ? p.flush_stages : k_iters;
  const int n_chunks = (k_iters + flush - 1) / flush;

  if (warp == 0 && lane == 0) {
    umma::tma_prefetch_desc(&tmA);
    umma::tma_prefetch_desc(&tmB);
    if (TS) umma::tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      umma::mbar_init(&full_bar[s], 1);
      umma::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      umma::mbar_init(&acc_full_bar[b], 1);
      umma::mbar_init(&acc_empty_bar[b], kConvEpiWarps * CG);    // CG == 2: both CTAs' epilogues report to the leader
    }
    umma::mbar_init(&cross_empty_bar[0], kConvEpiWarps * CG);
    umma::mbar_init(&cross_empty_bar[1], kConvEpiWarps * CG);
    umma::mbar_init(bres_bar, 1);
    umma::fence_barrier_init();
  } else if (warp == 1) {
    if (CG == 1) umma::tmem_alloc(tmem_ptr_smem, C::kTmemCols);
  }
  for (int i = threadIdx.x; i < p.Cout && i < C::kMaxBias; i += kConvThreads) bias_smem[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  umma::tc_fence_before();
  __syncthreads();
  if (CG == 2) {
    umma::cluster_sync();       // peer barriers are initialised before anything signals them
    // tcgen05.alloc.cta_group::2 is a compiler-generated handshake through the PEER CTA's reserved shared memory (remote
    // mbarrier arrive + remote store of the address): it may only run once the peer CTA is known to be executing, i.e. after a
    // cluster barrier.  Allocating before it hangs when the two CTAs of a pair start far apart, which several streams' kernels
    // sharing the GPU provoke (profiles/r02_two_cta_alloc_hang.txt).
    if (warp == 1) umma::tmem_alloc_2cta(tmem_ptr_smem, C::kTmemCols);
    umma::tc_fence_before();
    __syncthreads();
  }
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch, bias = a weight) overlapped
  // the tail of the previous kernel in the stream; from here on its outputs are read and its inputs overwritten
  pdl_wait();
  pdl_launch_dependents();

  // tile decode: n-tile fastest so CTAs that run concurrently share A tiles in L2
  // (for CG == 2 `t` indexes pairs of M tiles; this CTA takes M tile 2*pair + rank)
  auto decode = [&](int t, int& g, int& x0, int& y0, int& n0) {
    const int nt = t % p.n_tiles_n; t /= p.n_tiles_n;
    const int m_tiles = p.tiles_x * p.tiles_y;
    int m = CG == 2 ? (t % (m_tiles >> 1)) * 2 + (int)cta_rank : t % m_tiles;
    g = CG == 2 ? t / (m_tiles >> 1) : t / m_tiles;
    const int tx = m % p.tiles_x, ty = m / p.tiles_x;
    x0 = tx * p.TW; y0 = ty * p.TH; n0 = nt * BN;
  };
  const int total_work = CG == 2 ? p.total_tiles >> 1 : p.total_tiles;

  if (warp == 0) {
    {
      // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
      uint32_t git = 0;
      if (WR > 0) {   // the whole weight tensor, once (all groups and tiles of this layer share it: one N tile)
        if (umma::elect_one()) {
          umma::mbar_arrive_expect_tx(bres_bar, (uint32_t)C::kBResBytes);
          for (int t = 0; t < WR; ++t) {
            const int tap = t / p.k_chunks, kc = t - tap * p.k_chunks;
#pragma unroll
            for (int pl = 0; pl < P; ++pl)
              umma::tma_load_3d(smem + C::kBResOffset + (t * P + pl) * C::kBBytes, &tmB, bres_bar,
                                tap * p.Cin + kc * kConvBK, 0, pl);
          }
        }
        __syncwarp();
      }
      for (int tile = worker; tile < total_work; tile += n_workers) {
        int g, x0, y0, n0;
        decode(tile, g, x0, y0, n0);
        const int cin0 = p.cin_off + g * p.cin_group_stride;
        const int bk0 = g * p.b_group_k_stride;
        const int nb0 = n0 + (int)cta_rank * C::kBRows;      // this CTA's share of the weight rows
        for (int it = 0; it < k_iters; ++it, ++git) {
          const int s = git % STAGES;
          const uint32_t ph = (git / STAGES) & 1;
          umma::mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t fb = CG == 2 ? umma::mapa_u32(umma::smem_u32(&full_bar[s]), 0) : 0u;
          uint8_t* a_dst = smem + s * C::kStageBytes;
          uint8_t* b_dst = a_dst + P * C::kABytes;
          if (!umma::elect_one()) continue;
          // CG == 2: both CTAs' bytes are counted on the LEADER's full barrier
          if (leader) umma::mbar_arrive_expect_tx(&full_bar[s], C::kTxBytes * CG);
          if (NX == 3) {
            const int kc = it / 3, ky = it - kc * 3;       // channel chunk outer, kernel row inner
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
              if (CG == 2) umma::tma_load_4d_2cta(a_dst + pl * C::kABytes, &tmA, fb, cin0 + kc * kConvBK, x0 - 1, y0 + ky - 1, pl);
              else umma::tma_load_4d(a_dst + pl * C::kABytes, &tmA, &full_bar[s], cin0 + kc * kConvBK, x0 - 1, y0 + ky - 1, pl);
            }
#pragma unroll
            for (int kx = 0; kx < (WR ? 0 : 3); ++kx)
#pragma unroll
              for (int pl = 0; pl < P; ++pl) {
                const int kk = bk0 + (ky * 3 + kx) * p.Cin + kc * kConvBK;
                if (CG == 2) umma::tma_load_3d_2cta(b_dst + (kx * P + pl) * C::kBBytes, &tmB, fb, kk, nb0, pl);
                else umma::tma_load_3d(b_dst + (kx * P + pl) * C::kBBytes, &tmB, &full_bar[s], kk, nb0, pl);
              }
          } else {
            const int tap = it / p.k_chunks, kc = it - tap * p.k_chunks;
            const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
              if (CG == 2) umma::tma_load_4d_2cta(a_dst + pl * C::kABytes, &tmA, fb, cin0 + kc * kConvBK,
                                                  x0 * p.stride + kx - p.pad, y0 * p.stride + ky - p.pad, pl);
              else umma::tma_load_4d(a_dst + pl * C::kABytes, &tmA, &full_bar[s], cin0 + kc * kConvBK,
                                     x0 * p.stride + kx - p.pad, y0 * p.stride + ky - p.pad, pl);
            }
#pragma unroll
            for (int pl = 0; pl < (WR ? 0 : P); ++pl) {
              const int kk = bk0 + tap * p.Cin + kc * kConvBK;
              if (CG == 2) umma::tma_load_3d_2cta(b_dst + pl * C::kBBytes, &tmB, fb, kk, nb0, pl);
              else umma::tma_load_3d(b_dst + pl * C::kBBytes, &tmB, &full_bar[s], kk, nb0, pl);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only when CG == 2) =====================
      // the whole warp runs the loop; tcgen05.mma / commit are issued by one elected lane
      // kind::f16 format codes: 0 = fp16 (split mode, both planes), 1 = bf16 (single-plane mode)
      constexpr uint32_t kFmt = P == 2 ? 0u : 1u;
      constexpr uint32_t idesc = umma::idesc_f16kind_f32(kConvBM * CG, BN, kFmt, kFmt);
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (CG == 2) umma::mma_bf16_ss_2cta(d, a, b, id, acc);
        else umma::mma_bf16_ss(d, a, b, id, acc);
      };
      auto commit = [](uint64_t* bar) {
        if (CG == 2) umma::mma_commit_2cta(bar);
        else umma::mma_commit(bar);
      };
      uint32_t git = 0, gch = 0, tcount = 0;
      if (WR > 0) { umma::mbar_wait(bres_bar, 0); umma::tc_fence_after(); }
      const uint32_t bres_addr = umma::smem_u32(smem + C::kBResOffset);
      for (int tile = worker; tile < total_work; tile += n_workers, ++tcount) {
        const uint32_t xb = C::kCrossBufs == 2 ? (tcount & 1) : 0u;
        const uint32_t tmem_cross = tmem_base + (2 + xb) * BN;
        long long* trace = (p.dbg && tcount < 32) ? p.dbg + ((size_t)blockIdx.x * 32 + tcount) * 8 : nullptr;
        if (trace && lane == 0) trace[0] = clock64();
        if (P == 2) {   // the epilogue must have read this cross accumulator's previous use
          umma::mbar_wait(&cross_empty_bar[xb], (C::kCrossBufs == 2 ? ((tcount >> 1) & 1) : (tcount & 1)) ^ 1);
          umma::tc_fence_after();
        }
        if (trace && lane == 0) trace[1] = clock64();
        int it = 0;
        for (int chunk = 0; chunk < n_chunks; ++chunk, ++gch) {
          const int buf = gch & 1;
          const uint32_t tmem_main = tmem_base + buf * BN;
          umma::mbar_wait(&acc_empty_bar[buf], ((gch >> 1) & 1) ^ 1);   // drained two chunks ago
          umma::tc_fence_after();
          const int it_begin = it, it_end = min(it + flush, k_iters);
          for (; it < it_end; ++it, ++git) {
            const int s = git % STAGES;
            const uint32_t ph = (git / STAGES) & 1;
            umma::mbar_wait(&full_bar[s], ph);
            umma::tc_fence_after();
            const uint32_t a_addr = umma::smem_u32(smem + s * C::kStageBytes);
            uint32_t b_addr = a_addr + P * C::kABytes;
            if (WR > 0) {   // resident tile index t = tap * k_chunks + kc; the kx taps of a haloed stage are k_chunks apart
              int t0;
              if (NX == 3) { const int kc = it / 3, ky = it - kc * 3; t0 = ky * 3 * p.k_chunks + kc; }
              else t0 = it;                                  // it = tap * k_chunks + kc already
              b_addr = bres_addr + (uint32_t)(t0 * P * C::kBBytes);
            }
            const uint32_t b_tap_stride = WR > 0 ? (uint32_t)(p.k_chunks * P * C::kBBytes) : (uint32_t)(P * C::kBBytes);
            if (umma::elect_one()) {
#pragma unroll
            for (int kx = 0; kx < NX; ++kx) {
              // halo mode: tap kx reads rows [kx, kx+128) of the haloed A tile
              const uint64_t a_hi = umma::smem_desc_kmajor<kConvRowB>(a_addr + kx * kConvRowB);
              const uint64_t a_lo = umma::smem_desc_kmajor<kConvRowB>(a_addr + (P - 1) * C::kABytes + kx * kConvRowB);
              const uint64_t b_hi = umma::smem_desc_kmajor<kConvRowB>(b_addr + kx * b_tap_stride);
              const uint64_t b_lo = umma::smem_desc_kmajor<kConvRowB>(b_addr + kx * b_tap_stride + (P - 1) * C::kBBytes);
#pragma unroll
              for (int k = 0; k < kConvBK / 16; ++k) {
                const uint64_t koff = (uint64_t)(k * 32 >> 4);  // 16 elements = 32 bytes along K
                mma(tmem_main, a_hi + koff, b_hi + koff, idesc, (it != it_begin || kx != 0 || k != 0) ? 1u : 0u);
                if (P == 2) {   // hi*lo + lo*hi (lo*lo ~ 2^-22 relative is dropped)
                  mma(tmem_cross, a_hi + koff, b_lo + koff, idesc, (it | kx | k) != 0 ? 1u : 0u);
                  mma(tmem_cross, a_lo + koff, b_hi + koff, idesc, 1u);
                }
              }
            }
            commit(&empty_bar[s]);   // frees the smem stage (in both CTAs) once these MMAs have read it
            }
            __syncwarp();
          }
          if (umma::elect_one()) commit(&acc_full_bar[buf]);  // this chunk's accumulator (and all earlier MMAs) done
          __syncwarp();
          if (trace && lane == 0) trace[2 + (chunk == n_chunks - 1 ? 1 : 0)] = clock64();   // [2] first chunk issued, [3] all issued
        }
      }
    }
  } else {
    // ===================== epilogue (8 warps: 2 per TMEM lane quadrant) =====================
    if constexpr (TS) {
      // ---- streaming epilogue with TMA stores (one accumulation chunk per tile, see ConvCfg)
      const int q = warp & 3, hh = (warp - 2) >> 2, row = q * 32 + lane;
      const bool issuer = warp == 2 && lane == 0;
      uint8_t* stg0 = smem + C::kOutOffset;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      const uint32_t ae0 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&acc_empty_bar[0]), 0) : 0u;
      const uint32_t ae1 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&acc_empty_bar[1]), 0) : 0u;
      const uint32_t ce0 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&cross_empty_bar[0]), 0) : 0u;
      const uint32_t ce1 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&cross_empty_bar[1]), 0) : 0u;
      uint32_t gch = 0, tcount = 0, npass = 0;
      for (int tile = worker; tile < total_work; tile += n_workers, ++tcount, ++gch) {
        int g, x0, y0, n0;
        decode(tile, g, x0, y0, n0);
        const int buf = gch & 1;
        const uint32_t xb = C::kCrossBufs == 2 ? (tcount & 1) : 0u;
        umma::mbar_wait(&acc_full_bar[buf], (gch >> 1) & 1);
        umma::tc_fence_after();
        const uint32_t t_main = tmem_base + lane_base + (uint32_t)(buf * BN);
        const uint32_t t_cross = tmem_base + lane_base + (uint32_t)((2 + xb) * BN);
        const int py = y0 + row / p.TW, px = x0 + row % p.TW;
        const long long pix = (long long)py * p.W_out + px;
        const int ch_tile = p.cout_off + g * p.cout_group_stride + n0;
#pragma unroll 1
        for (int pass = 0; pass < BN / 64; ++pass, ++npass) {
          uint8_t* stg = stg0 + (npass & 1u) * 32768u;          // two staging tiles: a store stays in flight while the next is filled
          if (issuer) umma::tma_store_wait_read_but1();         // the store issued two passes ago has read this tile
          asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            const int col = pass * 64 + hh * 32 + g2 * 16;      // column of the BN-wide tile
            uint32_t m[16], c[16];
            umma::tmem_ld_32x16(t_main + (uint32_t)col, m);
            umma::tmem_ld_32x16(t_cross + (uint32_t)col, c);
            umma::tmem_ld_wait();
            float v[16];
            const float* bs = bias_smem + n0 + col;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = __fmaf_rn(__uint_as_float(m[j]) + __uint_as_float(c[j]), p.acc_scale, bs[j]);
              if (p.act == 1) v[j] = gelu_erf(x);
              else if (p.act == 4) v[j] = fmaxf(x, 0.f);
              else if (p.act == 2) v[j] = fast_sigmoid(x);
              else if (p.act == 3) v[j] = fast_tanh(x);
              else v[j] = x;
            }
            if (p.mask_src) {   // ReLU backward: pass the gradient where the forward activation was > 0
              const __nv_bfloat16* mk = p.mask_src + pix * p.Cout_total + ch_tile + col;
              uint32_t mb[8];
              *(uint4*)&mb[0] = *(const uint4*)mk;
              *(uint4*)&mb[4] = *(const uint4*)(mk + 8);
              if (p.mask_planes == 2) {
                uint32_t m2[8];
                *(uint4*)&m2[0] = *(const uint4*)(mk + p.mask_plane_stride);
                *(uint4*)&m2[4] = *(const uint4*)(mk + p.mask_plane_stride + 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) mb[j] |= m2[j];
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if ((mb[j] & 0x00007fffu) == 0u) v[2 * j] = 0.f;
                if ((mb[j] & 0x7fff0000u) == 0u) v[2 * j + 1] = 0.f;
              }
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) umma::pack_split2(v[2 * j], v[2 * j + 1], true, hi[j], lo[j]);
            // staging tile: [plane][128 rows][128 B], 16-byte unit u of row r lives at u ^ (r & 7)  (SWIZZLE_128B)
            const int u0 = hh * 4 + g2 * 2;
            uint8_t* r0 = stg + row * 128;
            *(uint4*)(r0 + (((u0) ^ (row & 7)) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *(uint4*)(r0 + (((u0 + 1) ^ (row & 7)) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *(uint4*)(r0 + 16384 + (((u0) ^ (row & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *(uint4*)(r0 + 16384 + (((u0 + 1) ^ (row & 7)) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
          if (pass == BN / 64 - 1) {      // the accumulators of this tile have been read: the MMA warp may reuse them
            umma::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) { umma::mbar_arrive_cluster(buf ? ae1 : ae0); umma::mbar_arrive_cluster(xb ? ce1 : ce0); }
              else { umma::mbar_arrive(&acc_empty_bar[buf]); umma::mbar_arrive(&cross_empty_bar[xb]); }
            }
          }
          umma::fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (issuer) {
            umma::tma_store_4d(&tmO, stg, ch_tile + pass * 64, x0, y0, 0);
            umma::tma_store_4d(&tmO, stg + 16384, ch_tile + pass * 64, x0, y0, 1);
            umma::tma_store_commit();
          }
        }
      }
      if (issuer) umma::tma_store_wait_all();
    } else {
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                 // which half of the BN columns
    const int row = q * 32 + lane;                    // tile row = output pixel within the tile
    const uint32_t lane_col = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * kHalf);
    uint32_t gch = 0;
    // CG == 2: "accumulator drained" is reported to the leader CTA, whose MMA warp owns the schedule
    const uint32_t ae0 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&acc_empty_bar[0]), 0) : 0u;
    const uint32_t ae1 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&acc_empty_bar[1]), 0) : 0u;
    const uint32_t ce0 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&cross_empty_bar[0]), 0) : 0u;
    const uint32_t ce1 = CG == 2 ? umma::mapa_u32(umma::smem_u32(&cross_empty_bar[1]), 0) : 0u;
    uint32_t tcount = 0;
    for (int tile = worker; tile < total_work; tile += n_workers, ++tcount) {
      int g, x0, y0, n0;
      decode(tile, g, x0, y0, n0);
      float acc[kHalf];
#pragma unroll
      for (int j = 0; j < kHalf; ++j) acc[j] = 0.f;
      long long* trace = (p.dbg && tcount < 32 && warp == 2) ? p.dbg + ((size_t)blockIdx.x * 32 + tcount) * 8 : nullptr;
      if (trace && lane == 0) trace[4] = clock64();                     // epilogue ready for this tile
      for (int chunk = 0; chunk < n_chunks; ++chunk, ++gch) {
        const int buf = gch & 1;
        umma::mbar_wait(&acc_full_bar[buf], (gch >> 1) & 1);
        umma::tc_fence_after();
        if (trace && lane == 0 && chunk == n_chunks - 1) trace[5] = clock64();   // last chunk's MMAs complete
        auto drain = [&](uint32_t col0) {   // acc += TMEM[col0 .. col0 + kHalf): every 16-column load in flight before ONE wait
          uint32_t r[kHalf];                // (a wait per pair of loads cost ~700 cycles each: 2900 cycles for the final drain of a 128-wide tile)
#pragma unroll
          for (int gi = 0; gi < kGroups; ++gi) umma::tmem_ld_32x16(tmem_base + lane_col + col0 + (uint32_t)(gi * 16), r + gi * 16);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < kHalf; ++j) acc[j] += __uint_as_float(r[j]);
        };
        drain((uint32_t)(buf * BN));
        const uint32_t xb = C::kCrossBufs == 2 ? (tcount & 1) : 0u;
        if (P == 2 && chunk == n_chunks - 1) drain((uint32_t)((2 + xb) * BN));   // the last commit also covers every cross-term MMA
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) {
            umma::mbar_arrive_cluster(buf ? ae1 : ae0);
            if (P == 2 && chunk == n_chunks - 1) umma::mbar_arrive_cluster(xb ? ce1 : ce0);
          } else {
            umma::mbar_arrive(&acc_empty_bar[buf]);
            if (P == 2 && chunk == n_chunks - 1) umma::mbar_arrive(&cross_empty_bar[xb]);
          }
        }
      }
      if (trace && lane == 0) trace[6] = clock64();                     // accumulators drained
      // ---- bias / activation / store (overlaps the next tile's main loop).  The activation is dispatched
      // ONCE per tile to a specialised instance: a per-element runtime select made the compiler evaluate
      // erff, expf, tanhf ... for every output (measured: ~9000 instructions per warp per tile).
      const int py = y0 + row / p.TW, px = x0 + row % p.TW;
      const long long pix = (long long)py * p.W_out + px + (long long)g * p.out_group_pix_stride;
      const long long ch0 = (long long)p.cout_off + (long long)g * p.cout_group_stride + n0 + half * kHalf;
      const float* bias = p.bias ? p.bias + n0 + half * kHalf : nullptr;
      if (p.act < 5 && p.Cout <= C::kMaxBias) {
        const float* bias_s = bias_smem + n0 + half * kHalf;
        if (p.border_bias) {
          const int yc = py == 0 ? 0 : (py == p.H_out - 1 ? 2 : 1), xc = px == 0 ? 0 : (px == p.W_out - 1 ? 2 : 1);
          if (yc != 1 || xc != 1) bias_s = p.border_bias + (size_t)(yc * 3 + xc) * p.Cout + n0 + half * kHalf;
        }
        switch (p.act) {
          case 0: conv_epilogue_fast<0, kHalf>(acc, p, pix, ch0, bias_s); break;
          case 1: conv_epilogue_fast<1, kHalf>(acc, p, pix, ch0, bias_s); break;
          case 2: conv_epilogue_fast<2, kHalf>(acc, p, pix, ch0, bias_s); break;
          case 3: conv_epilogue_fast<3, kHalf>(acc, p, pix, ch0, bias_s); break;
          default: conv_epilogue_fast<4, kHalf>(acc, p, pix, ch0, bias_s); break;
        }
        if (trace && lane == 0) trace[7] = clock64();                   // tile stored
        continue;
      }
      switch (p.act) {
        case 0: conv_epilogue<0, kHalf>(acc, p, pix, ch0, bias); break;
        case 1: conv_epilogue<1, kHalf>(acc, p, pix, ch0, bias); break;
        case 2: conv_epilogue<2, kHalf>(acc, p, pix, ch0, bias); break;
        case 3: conv_epilogue<3, kHalf>(acc, p, pix, ch0, bias); break;
        case 4: conv_epilogue<4, kHalf>(acc, p, pix, ch0, bias); break;
        case 5: conv_epilogue<5, kHalf>(acc, p, pix, ch0, bias); break;
        default: conv_epilogue<6, kHalf>(acc, p, pix, ch0, bias); break;
      }
    }
    }   // !TS
  }
  umma::tc_fence_before();
  __syncthreads();
  if (CG == 2) umma::cluster_sync();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    if (CG == 2) umma::tmem_dealloc_2cta(tmem_base, C::kTmemCols);
    else umma::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ wide tiles for the 256-channel encoder layers
// k_conv_wide: 256 output pixels x 256 output channels per CTA pair (cta_group::2, N = 256), split planes,
// one tap per stage.  The 256->256 layers on 64x64 images were bound by L2->SM operand traffic: with BN = 128
// every activation tile was fetched once per N tile (64 B/cycle/SM needed, ~42 available).  One N tile halves
// the activation traffic.  The two accumulators (hi*hi, cross) take all 512 TMEM columns, so there is no
// ping-pong: the chain is not cut (K/16 <= 144 hi*hi MMAs; the measured flow error stays two orders below the
// 1e-4 budget) and the epilogue streams TMEM -> activation -> store 16 columns at a time.
// BK = 64 channels per stage here: 128-byte rows (SWIZZLE_128B), i.e. full-line L2 requests instead of half lines
constexpr int kWideBK = 64;
constexpr int kWideRowB = 128;
constexpr int kWideStageBytes = 2 * 128 * kWideRowB + 2 * 128 * kWideRowB;   // A: 2 planes x 128 px, B: 2 planes x 128 rows
constexpr int kWideStages = 3;
constexpr int kWideBarOffset = kWideStages * kWideStageBytes;                // 196608
constexpr int kWideBiasOffset = kWideBarOffset + 256;
constexpr int kWideTotal = kWideBiasOffset + 256 * 4 + 1024;    // + 1024-byte alignment slack

__global__ void __launch_bounds__(kConvThreads, 1)
k_conv_wide(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kWideBarOffset);
  uint64_t* empty_bar = full_bar + kWideStages;
  uint64_t* acc_full_bar = empty_bar + kWideStages;
  uint64_t* acc_empty_bar = acc_full_bar + 1;
  uint32_t* tmem_ptr_smem = (uint32_t*)(acc_empty_bar + 1);
  float* bias_smem = (float*)(smem + kWideBiasOffset);
  constexpr int kABytes = 128 * kWideRowB, kBBytes = 128 * kWideRowB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = umma::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_workers = (int)(gridDim.x >> 1), worker = (int)(blockIdx.x >> 1);
  const int k_chunks = p.Cin / kWideBK;
  const int k_iters = p.taps * k_chunks;
  if (warp == 0 && lane == 0) {
    umma::tma_prefetch_desc(&tmA);
    umma::tma_prefetch_desc(&tmB);
    for (int s = 0; s < kWideStages; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
    umma::mbar_init(acc_full_bar, 1);
    umma::mbar_init(acc_empty_bar, kConvEpiWarps * 2);
    umma::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += kConvThreads) bias_smem[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  __syncthreads();
  umma::cluster_sync();
  // tcgen05.alloc.cta_group::2 is a compiler-generated handshake through the PEER CTA's reserved shared memory (remote
  // mbarrier arrive + remote store of the address): it may only run once the peer CTA is known to be executing, i.e. after a
  // cluster barrier.  Allocating before it hangs when the two CTAs of a pair start far apart, which several streams' kernels
  // sharing the GPU provoke (profiles/r02_two_cta_alloc_hang.txt).
  if (warp == 1) umma::tmem_alloc_2cta(tmem_ptr_smem, 512);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();
  const int m_tiles = p.tiles_x * p.tiles_y;
  const int total_work = (m_tiles >> 1) * p.n_groups;
  auto decode = [&](int t, int& g, int& x0, int& y0) {
    const int m = (t % (m_tiles >> 1)) * 2 + (int)cta_rank;
    g = t / (m_tiles >> 1);
    x0 = (m % p.tiles_x) * p.TW; y0 = (m / p.tiles_x) * p.TH;
  };

  if (warp == 0) {
    uint32_t git = 0;
    for (int tile = worker; tile < total_work; tile += n_workers) {
      int g, x0, y0;
      decode(tile, g, x0, y0);
      const int cin0 = p.cin_off + g * p.cin_group_stride;
      for (int it = 0; it < k_iters; ++it, ++git) {
        const int s = git % kWideStages;
        umma::mbar_wait(&empty_bar[s], ((git / kWideStages) & 1) ^ 1);
        const uint32_t fb = umma::mapa_u32(umma::smem_u32(&full_bar[s]), 0);
        uint8_t* a_dst = smem + s * kWideStageBytes;
        uint8_t* b_dst = a_dst + 2 * kABytes;
        if (umma::elect_one()) {
          if (leader) umma::mbar_arrive_expect_tx(&full_bar[s], kWideStageBytes * 2);
          const int tap = it / k_chunks, kc = it - tap * k_chunks;
          const int ky = tap / p.ksize, kx = tap - ky * p.ksize;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
            umma::tma_load_4d_2cta(a_dst + pl * kABytes, &tmA, fb, cin0 + kc * kWideBK, x0 * p.stride + kx - p.pad,
                                   y0 * p.stride + ky - p.pad, pl);
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
            umma::tma_load_3d_2cta(b_dst + pl * kBBytes, &tmB, fb, tap * p.Cin + kc * kWideBK, (int)cta_rank * 128, pl);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma::idesc_f16kind_f32(256, 256, 0u, 0u);
      uint32_t git = 0, tcount = 0;
      for (int tile = worker; tile < total_work; tile += n_workers, ++tcount) {
        umma::mbar_wait(acc_empty_bar, (tcount & 1) ^ 1);     // both CTAs' epilogues have streamed the previous tile out
        umma::tc_fence_after();
        for (int it = 0; it < k_iters; ++it, ++git) {
          const int s = git % kWideStages;
          umma::mbar_wait(&full_bar[s], (git / kWideStages) & 1);
          umma::tc_fence_after();
          const uint32_t a_addr = umma::smem_u32(smem + s * kWideStageBytes), b_addr = a_addr + 2 * kABytes;
          if (umma::elect_one()) {
            const uint64_t a_hi = umma::smem_desc_kmajor<kWideRowB>(a_addr), a_lo = umma::smem_desc_kmajor<kWideRowB>(a_addr + kABytes);
            const uint64_t b_hi = umma::smem_desc_kmajor<kWideRowB>(b_addr), b_lo = umma::smem_desc_kmajor<kWideRowB>(b_addr + kBBytes);
#pragma unroll
            for (int k = 0; k < kWideBK / 16; ++k) {
              const uint64_t koff = (uint64_t)(k * 32 >> 4);
              umma::mma_bf16_ss_2cta(tmem_base, a_hi + koff, b_hi + koff, idesc, (it | k) != 0 ? 1u : 0u);
              umma::mma_bf16_ss_2cta(tmem_base + 256, a_hi + koff, b_lo + koff, idesc, (it | k) != 0 ? 1u : 0u);
              umma::mma_bf16_ss_2cta(tmem_base + 256, a_lo + koff, b_hi + koff, idesc, 1u);
            }
            umma::mma_commit_2cta(&empty_bar[s]);
          }
          __syncwarp();
        }
        if (umma::elect_one()) umma::mma_commit_2cta(acc_full_bar);
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t lane_col = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
    const uint32_t ae = umma::mapa_u32(umma::smem_u32(acc_empty_bar), 0);
    uint32_t tcount = 0;
    for (int tile = worker; tile < total_work; tile += n_workers, ++tcount) {
      int g, x0, y0;
      decode(tile, g, x0, y0);
      const int py = y0 + row / p.TW, px = x0 + row % p.TW;
      const long long pix = (long long)py * p.W_out + px;
      const long long ch0 = (long long)p.cout_off + (long long)g * p.cout_group_stride + half * 128;
      EpiOut o;
      o.scale = p.acc_scale;
      o.mode = p.out_fp32 ? 0 : (p.out_planes == 2 ? 2 : 1);
      o.plane_stride_b = p.out_plane_stride * 2;
      o.mask = nullptr; o.mask_plane_stride = 0; o.mask_planes = 0;
      o.out_t = nullptr; o.out_t_plane_stride = 0; o.ld_t = 0; o.t_planes = 1;
      const long long elem = pix * p.Cout_total + ch0;
      umma::mbar_wait(acc_full_bar, tcount & 1);
      umma::tc_fence_after();
#pragma unroll 1
      for (int gi = 0; gi < 8; ++gi) {
        uint32_t m[16], c[16];
        umma::tmem_ld_32x16(tmem_base + lane_col + (uint32_t)(gi * 16), m);
        umma::tmem_ld_32x16(tmem_base + lane_col + (uint32_t)(256 + gi * 16), c);
        umma::tmem_ld_wait();
        float a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = __uint_as_float(m[j]) + __uint_as_float(c[j]);
        o.out = p.out_fp32 ? (void*)((float*)p.out + elem + gi * 16) : (void*)((__nv_bfloat16*)p.out + elem + gi * 16);
        const float* bs = bias_smem + half * 128 + gi * 16;
        if (p.act == 1)
          epi_store16<1>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15], bs, o);
        else
          epi_store16<0>(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15], bs, o);
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive_cluster(ae);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync();
  if (warp == 1) umma::tmem_dealloc_2cta(tmem_base, 512);
}

// ------------------------------------------------------------------ two output rows per CTA (decoder-half 3x3 layers)
// k_conv_rows2: 3x3 stride-1 convolution, split planes, BN = 96, rows of >= 256 pixels.  A CTA pair computes TWO output
// rows x 256 pixels x 96 channels.  Why: with one output row per CTA the u4/u5 layers are bound by the rate at which an
// SM ingests operands from L2 (~28 B/cycle/SM, profiles/r01_conv_fill_rate_model.txt): per 32-channel chunk a CTA
// loads 3 haloed activation rows + the 9 weight taps for 54 MMAs (35-40 B per MMA cycle).  Two output rows share
// their activation rows (4 haloed rows instead of 6) and every weight tap (loaded once for both rows), so the same
// chunk costs 4 rows + 9 taps for 108 MMAs: ~20 B per MMA cycle, below the fill rate -- the tensor pipe becomes the bound.
// Pipeline item = (chunk kc, activation row a in 0..3): TMA lands haloed row a (130 px x 32 ch x 2 planes) and, for
// a < 3, the three kx taps of kernel row ky = a (both planes, this CTA's half of the 96 weight rows).  Item a feeds
// output row 0 with ky = a and output row 1 with ky = a - 1, so the weights of item a-1 are used once more: a slot
// is released one item late.  TMEM: main and cross accumulators of both rows (4 x 96 columns); the hi*hi chain is
// NOT cut here (K/16 <= 256 MMAs per chain is enforced by the host; the layers that use this kernel have 54-216),
// so the epilogue is a single drain per row followed by bias / activation / split / store.
constexpr int kR2BN = 96;
constexpr int kR2ABytes = 136 * kConvRowB;                 // one haloed row, one plane (TMA writes 130 rows)
constexpr int kR2ATx = 130 * kConvRowB;
constexpr int kR2BBytes = (kR2BN / 2) * kConvRowB;         // one tap, one plane, this CTA's 48 weight rows
constexpr int kR2StageBytes = 2 * kR2ABytes + 6 * kR2BBytes;   // 17408 + 18432
constexpr int kR2Stages = 6;
constexpr int kR2BarOffset = kR2Stages * kR2StageBytes;    // 215040
constexpr int kR2BiasOffset = kR2BarOffset + 256;
constexpr int kR2MaxBias = 1024;
constexpr int kR2Total = kR2BiasOffset + kR2MaxBias * 4 + 1024;

template <int ACT>
__device__ __forceinline__ void rows2_store(const float* acc, const ConvParams& p, int py, int px, long long ch0,
                                            const float* bias_smem, int nofs) {
  const long long pix = (long long)py * p.W_out + px;
  const float* bias_s = bias_smem + nofs;
  if (p.border_bias) {
    const int yc = py == 0 ? 0 : (py == p.H_out - 1 ? 2 : 1), xc = px == 0 ? 0 : (px == p.W_out - 1 ? 2 : 1);
    if (yc != 1 || xc != 1) bias_s = p.border_bias + (size_t)(yc * 3 + xc) * p.Cout + nofs;
  }
  conv_epilogue_fast<ACT, kR2BN / 2>(acc, p, pix, ch0, bias_s);
}

__global__ void __launch_bounds__(kConvThreads, 1)
k_conv_rows2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kR2BarOffset);
  uint64_t* empty_bar = full_bar + kR2Stages;
  uint64_t* acc_full_bar = empty_bar + kR2Stages;     // [2] one per output row
  uint64_t* acc_empty_bar = acc_full_bar + 2;         // [2]
  uint32_t* tmem_ptr_smem = (uint32_t*)(acc_empty_bar + 2);
  float* bias_smem = (float*)(smem + kR2BiasOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = umma::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_workers = (int)(gridDim.x >> 1), worker = (int)(blockIdx.x >> 1);
  const int KC = p.k_chunks;
  if (warp == 0 && lane == 0) {
    umma::tma_prefetch_desc(&tmA);
    umma::tma_prefetch_desc(&tmB);
    for (int s = 0; s < kR2Stages; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
    for (int r = 0; r < 2; ++r) { umma::mbar_init(&acc_full_bar[r], 1); umma::mbar_init(&acc_empty_bar[r], kConvEpiWarps * 2); }
    umma::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.Cout && i < kR2MaxBias; i += kConvThreads) bias_smem[i] = p.bias ? __ldg(p.bias + i) : 0.f;
  __syncthreads();
  umma::cluster_sync();
  // tcgen05.alloc.cta_group::2 is a compiler-generated handshake through the PEER CTA's reserved shared memory (remote
  // mbarrier arrive + remote store of the address): it may only run once the peer CTA is known to be executing, i.e. after a
  // cluster barrier.  Allocating before it hangs when the two CTAs of a pair start far apart, which several streams' kernels
  // sharing the GPU provoke (profiles/r02_two_cta_alloc_hang.txt).
  if (warp == 1) umma::tmem_alloc_2cta(tmem_ptr_smem, 512);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  // pair tile t: n-tile fastest (pairs that run concurrently share activation rows in L2), then pair column, then row pair
  const int pair_cols = p.W_out >> 8;
  const int total_work = pair_cols * (p.H_out >> 1) * p.n_tiles_n;
  auto decode = [&](int t, int& x0, int& y0, int& n0) {
    const int nt = t % p.n_tiles_n; t /= p.n_tiles_n;
    const int pc = t % pair_cols;
    x0 = (pc * 2 + (int)cta_rank) * 128; y0 = (t / pair_cols) * 2; n0 = nt * kR2BN;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t git = 0;
    for (int tile = worker; tile < total_work; tile += n_workers) {
      int x0, y0, n0;
      decode(tile, x0, y0, n0);
      const int nb0 = n0 + (int)cta_rank * (kR2BN / 2);
      for (int kc = 0; kc < KC; ++kc) {
        for (int a = 0; a < 4; ++a, ++git) {
          const int s = git % kR2Stages;
          umma::mbar_wait(&empty_bar[s], ((git / kR2Stages) & 1) ^ 1);
          const uint32_t fb = umma::mapa_u32(umma::smem_u32(&full_bar[s]), 0);
          uint8_t* a_dst = smem + s * kR2StageBytes;
          uint8_t* b_dst = a_dst + 2 * kR2ABytes;
          if (umma::elect_one()) {
            if (leader) umma::mbar_arrive_expect_tx(&full_bar[s], 2u * (2u * kR2ATx + (a < 3 ? 6u * kR2BBytes : 0u)));
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
              umma::tma_load_4d_2cta(a_dst + pl * kR2ABytes, &tmA, fb, p.cin_off + kc * kConvBK, x0 - 1, y0 + a - 1, pl);
            if (a < 3) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
                  umma::tma_load_3d_2cta(b_dst + (kx * 2 + pl) * kR2BBytes, &tmB, fb, (a * 3 + kx) * p.Cin + kc * kConvBK, nb0, pl);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = umma::idesc_f16kind_f32(256, kR2BN, 0u, 0u);
      uint32_t git = 0, tcount = 0;
      for (int tile = worker; tile < total_work; tile += n_workers, ++tcount) {
        for (int kc = 0; kc < KC; ++kc) {
          for (int a = 0; a < 4; ++a, ++git) {
            const int s = git % kR2Stages;
            if (kc == 0 && a < 2) {   // first MMAs into output row `a` of this tile: the epilogue has drained the previous tile's
              umma::mbar_wait(&acc_empty_bar[a], (tcount & 1) ^ 1);
              umma::tc_fence_after();
            }
            umma::mbar_wait(&full_bar[s], (git / kR2Stages) & 1);
            umma::tc_fence_after();
            const uint32_t a_addr = umma::smem_u32(smem + s * kR2StageBytes);
            if (umma::elect_one()) {
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const int ky = a - r;
                if (ky < 0 || ky > 2) continue;
                // weights of kernel row ky live in the slot of item (kc, ky) = git - a + ky
                const uint32_t b_addr = umma::smem_u32(smem + ((git - (uint32_t)a + (uint32_t)ky) % kR2Stages) * kR2StageBytes) + 2 * kR2ABytes;
                const uint32_t t_main = tmem_base + (uint32_t)(r * kR2BN), t_cross = tmem_base + (uint32_t)((2 + r) * kR2BN);
                const bool first = kc == 0 && ky == 0;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const uint64_t a_hi = umma::smem_desc_kmajor<kConvRowB>(a_addr + kx * kConvRowB);
                  const uint64_t a_lo = umma::smem_desc_kmajor<kConvRowB>(a_addr + kR2ABytes + kx * kConvRowB);
                  const uint64_t b_hi = umma::smem_desc_kmajor<kConvRowB>(b_addr + (kx * 2) * kR2BBytes);
                  const uint64_t b_lo = umma::smem_desc_kmajor<kConvRowB>(b_addr + (kx * 2 + 1) * kR2BBytes);
#pragma unroll
                  for (int k = 0; k < kConvBK / 16; ++k) {
                    const uint64_t koff = (uint64_t)(k * 32 >> 4);
                    const uint32_t acc = (first && kx == 0 && k == 0) ? 0u : 1u;
                    umma::mma_bf16_ss_2cta(t_main, a_hi + koff, b_hi + koff, idesc, acc);
                    umma::mma_bf16_ss_2cta(t_cross, a_hi + koff, b_lo + koff, idesc, acc);
                    umma::mma_bf16_ss_2cta(t_cross, a_lo + koff, b_hi + koff, idesc, 1u);
                  }
                }
              }
              // item a-1 (its weights were used once more just now) and, at the end of a chunk, item 3 are free
              if (a >= 1) umma::mma_commit_2cta(&empty_bar[(git - 1) % kR2Stages]);
              if (a == 3) umma::mma_commit_2cta(&empty_bar[s]);
              if (kc == KC - 1 && a >= 2) umma::mma_commit_2cta(&acc_full_bar[a - 2]);   // row 0 ends with item 2, row 1 with item 3
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===================== epilogue: drain both rows, release TMEM, then bias / activation / split / store =====================
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    constexpr int kHalf = kR2BN / 2;
    const uint32_t lane_col = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * kHalf);
    const uint32_t ae0 = umma::mapa_u32(umma::smem_u32(&acc_empty_bar[0]), 0);
    const uint32_t ae1 = umma::mapa_u32(umma::smem_u32(&acc_empty_bar[1]), 0);
    uint32_t tcount = 0;
    for (int tile = worker; tile < total_work; tile += n_workers, ++tcount) {
      int x0, y0, n0;
      decode(tile, x0, y0, n0);
      float acc[2][kHalf];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        umma::mbar_wait(&acc_full_bar[r], tcount & 1);
        umma::tc_fence_after();
#pragma unroll
        for (int gi = 0; gi < kHalf / 16; ++gi) {
          uint32_t m[16], c[16];
          umma::tmem_ld_32x16(tmem_base + lane_col + (uint32_t)(r * kR2BN + gi * 16), m);
          umma::tmem_ld_32x16(tmem_base + lane_col + (uint32_t)((2 + r) * kR2BN + gi * 16), c);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[r][gi * 16 + j] = __uint_as_float(m[j]) + __uint_as_float(c[j]);
        }
        umma::tc_fence_before();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive_cluster(r ? ae1 : ae0);
      }
      const int px = x0 + row;
      const long long ch0 = (long long)p.cout_off + n0 + half * kHalf;
      const int nofs = n0 + half * kHalf;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        switch (p.act) {
          case 0: rows2_store<0>(acc[r], p, y0 + r, px, ch0, bias_smem, nofs); break;
          case 1: rows2_store<1>(acc[r], p, y0 + r, px, ch0, bias_smem, nofs); break;
          case 2: rows2_store<2>(acc[r], p, y0 + r, px, ch0, bias_smem, nofs); break;
          case 3: rows2_store<3>(acc[r], p, y0 + r, px, ch0, bias_smem, nofs); break;
          default: rows2_store<4>(acc[r], p, y0 + r, px, ch0, bias_smem, nofs); break;
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync();
  if (warp == 1) umma::tmem_dealloc_2cta(tmem_base, 512);
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
  pdl_wait();
  pdl_launch_dependents();
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

static int g_persistent = 1;
static int g_enable_2cta = 1;
static int g_weights_resident = 1;
static int g_wide_tiles = 1;
static int g_pair_min_mmas = 48;
static int g_max_sms = kNumSMs;   // experiment knob: SMs a persistent launch may occupy
// Streaming epilogue with TMA stores for short-K 128-wide tiles (ConvCfg TS).  Correct (tests/test_gpu_conv.py runs it), but
// measured SLOWER than the per-thread stores on the one shape it targets -- the FastNSF 128x128 GEMM: 34.3 -> 47.1 us with
// double-buffered staging tiles (profiles/r02_conv_tma_store_ab.txt).  Those GEMMs turned out to be bound by SM<->L2 bytes
// (128 KB of operands + 64 KB of output per 128-point tile at ~30 B/cycle/SM), not by the store instruction pattern, and the
// staging adds two block-wide barriers per 64 columns.  Off by default; kept as an A/B knob.
static int g_tma_store = 0;
static long long* g_dbg = nullptr;
static int g_rows2 = 1;    // two output rows per CTA pair for the 96-channel decoder-half layers (k_conv_rows2)
static int g_pdl = 1;      // programmatic dependent launch between the backbone's kernels

template <int BN, int P, int NX, int CG, int WR = 0, int TS = 0>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t stream,
                       const CUtensorMap* tmO = nullptr) {
  using C = ConvCfg<BN, P, NX, CG, WR, TS>;
  auto kern = k_conv_umma<BN, P, NX, CG, WR, TS>;
  static bool configured_dev[64] = {};      // the attribute is per device: one flag per device ordinal
  int dev_ = 0;
  HIMO_CUDA_RET(cudaGetDevice(&dev_));
  bool& configured = configured_dev[dev_ & 63];
  if (!configured) {
    HIMO_CUDA_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    configured = true;
  }
  const int work = p.total_tiles / CG;
  int grid = (work < g_max_sms / CG || !g_persistent) ? work : g_max_sms / CG;   // persistent: one CTA (pair) per SM (pair)
  grid *= CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = C::kTotal; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
  HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO ? *tmO : tmA, p));
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

}  // namespace himo

using namespace himo;

static int g_flush_mmas = 48;
static int g_disable_halo = 0;
// Debug / A-B knob: 1 disables the haloed-row reuse (every tap loads its own A tile).
extern "C" int himo_conv_set_halo(int enable) { g_disable_halo = enable ? 0 : 1; return HIMO_OK; }
// A/B knob: 0 disables the CTA-pair (cta_group::2) path.
extern "C" int himo_conv_set_2cta(int enable) { g_enable_2cta = enable ? 1 : 0; return HIMO_OK; }
// Retired experiment (round 1: the A operand staged in tensor memory, tcgen05.cp + TS-form MMAs -- bit-identical, neutral to
// slower, profiles/r01_conv_a_tmem_ab.txt).  Its 128 staging columns now hold the second cross accumulator of the 128-wide
// tiles; the entry point stays so that existing callers keep linking.
extern "C" int himo_conv_set_a_tmem(int) { return HIMO_OK; }
// Experiment knob: number of SMs a persistent convolution launch occupies (default all 148); used to tell a per-SM
// operand-fill limit from a chip-wide one.
extern "C" int himo_conv_set_max_sms(int n) { g_max_sms = n < 2 ? 2 : (n > kNumSMs ? kNumSMs : n); return HIMO_OK; }
// Tuning knob: CTA pairs are used when a tile carries at least this many hi*hi MMAs (default 48).
extern "C" int himo_conv_set_pair_min_mmas(int n) { g_pair_min_mmas = n; return HIMO_OK; }
// A/B knob: 0 disables the 256-wide N tiles (k_conv_wide) of the 256-channel encoder layers.
extern "C" int himo_conv_set_wide_tiles(int enable) { g_wide_tiles = enable ? 1 : 0; return HIMO_OK; }
// A/B knob: 0 disables the weights-resident variants of the 64-channel encoder layers.
extern "C" int himo_conv_set_weights_resident(int enable) { g_weights_resident = enable ? 1 : 0; return HIMO_OK; }
// A/B knob: 0 disables the TMA-store epilogue (every thread stores its own 16-byte pieces).
extern "C" int himo_conv_set_tma_store(int enable) { g_tma_store = enable ? 1 : 0; return HIMO_OK; }
// Profiling hook: device buffer of [grid][32][8] int64 that k_conv_umma fills with clock64() marks per tile (NULL = off).
extern "C" int himo_conv_set_debug_buffer(void* buf) { g_dbg = (long long*)buf; return HIMO_OK; }
// A/B knob: 0 disables the two-output-rows tiles (k_conv_rows2) of the 96-channel decoder-half layers.
extern "C" int himo_conv_set_rows2(int enable) { g_rows2 = enable ? 1 : 0; return HIMO_OK; }
// A/B knob: 0 launches the backbone kernels without programmatic dependent launch (full stream serialisation).
extern "C" int himo_conv_set_pdl(int enable) { g_pdl = enable ? 1 : 0; return HIMO_OK; }
// A/B knob: 0 launches one CTA per tile instead of the persistent one-CTA-per-SM tile loop.
extern "C" int himo_conv_set_persistent(int enable) { g_persistent = enable ? 1 : 0; return HIMO_OK; }
// Tuning knob (process-wide): number of hi*hi MMAs (K = 16 each) accumulated in tensor memory before the
// partial sum is drained into fp32 registers (split mode).  Kept under the historical name.
extern "C" int himo_conv_set_flush_iters(int mmas) {
  if (mmas < 1) return HIMO_ERR_ARG;
  g_flush_mmas = mmas;
  return HIMO_OK;
}

extern "C" int himo_conv2d_nhwc(const himo_conv_desc* d, void* stream_) {
  if (!d || !d->in || !d->wgt || (!d->out && d->act < 5)) return HIMO_ERR_ARG;
  if (d->act >= 5 && (!d->aux_h || !d->aux_z || !d->out2 || d->aux_ld % 16)) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  constexpr int BK = kConvBK;
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

  // CTA pairs (cta_group::2) whenever the M tiles pair up
  const int m_tiles_total = (W_out / TW) * (H_out / TH);
  // (measured: pairing pays once a tile carries >= ~48 hi*hi MMAs; below that the pair's lock-step costs more
  // than the halved weight traffic saves)
  const int main_mmas_per_tile = d->ksize * d->ksize * (d->Cin / BK) * 2;
  const int CGsel = (g_enable_2cta && m_tiles_total % 2 == 0 && main_mmas_per_tile >= g_pair_min_mmas) ? 2 : 1;
  // halo mode: 3x3, stride 1, full 128-pixel row tiles -> one haloed A load feeds the three kx taps
  const bool halo = d->ksize == 3 && d->stride == 1 && TW == 128 && TH == 1 && !g_disable_halo;
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) return HIMO_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->Cin_total, (cuuint64_t)d->W_in, (cuuint64_t)d->H_in, (cuuint64_t)P};
    cuuint64_t strides[3] = {(cuuint64_t)d->Cin_total * 2, (cuuint64_t)d->W_in * d->Cin_total * 2,
                             (cuuint64_t)d->in_plane_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(halo ? TW + 2 : (TW - 1) * d->stride + 1),
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
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(BN / CGsel), 1};
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
  p.W_out = W_out; p.Cout_total = d->Cout_total; p.cout_off = d->cout_off; p.Cout = d->Cout;
  p.cout_group_stride = d->cout_group_stride; p.act = d->act; p.out_fp32 = d->out_fp32;
  p.acc_scale = d->acc_scale != 0.f ? d->acc_scale : 1.f;
  {  // stages per accumulation chain: a stage issues 2*NX hi*hi MMAs
    const int per_stage = 2 * (halo ? 3 : 1);
    p.flush_stages = g_flush_mmas / per_stage > 0 ? g_flush_mmas / per_stage : 1;
  }
  p.out_t = (__nv_bfloat16*)d->out_t; p.out_t_plane_stride = d->out_t_plane_stride; p.ld_t = d->ld_t;
  p.mask_src = (const __nv_bfloat16*)d->mask_src; p.mask_plane_stride = d->mask_plane_stride;
  p.mask_planes = d->mask_planes;
  p.b_group_k_stride = d->b_group_k_stride; p.out_group_pix_stride = d->out_group_pix_stride;
  p.stop_flag = d->stop_flag;
  p.aux_h = d->aux_h; p.aux_z = d->aux_z; p.aux_ld = d->aux_ld;
  p.out2 = (__nv_bfloat16*)d->out2; p.out2_plane_stride = d->out2_plane_stride; p.out2_ld = d->out2_ld;
  p.border_bias = d->border_bias; p.H_out = H_out; p.dbg = g_dbg;
  if (d->border_bias && (d->act >= 5 || d->Cout > 1024)) return HIMO_ERR_UNSUPPORTED;
  p.total_tiles = p.tiles_x * p.tiles_y * p.n_tiles_n * groups;
  // two output rows per CTA pair for the decoder-half 3x3 layers with 96-channel N tiles (k_conv_rows2)
  if (g_rows2 && P == 2 && halo && BN == 96 && W_out % 256 == 0 && H_out % 2 == 0 && groups == 1 && d->act < 5 &&
      d->Cout <= kR2MaxBias && !d->out_t && !d->mask_src && !d->stop_flag && d->b_group_k_stride == 0 && !d->b_k_total &&
      taps * p.k_chunks * 2 <= 256 && CGsel == 2) {
    static bool r2_configured_dev[64] = {};
    int dev_ = 0;
    HIMO_CUDA_RET(cudaGetDevice(&dev_));
    if (!r2_configured_dev[dev_ & 63]) {
      HIMO_CUDA_RET(cudaFuncSetAttribute(k_conv_rows2, cudaFuncAttributeMaxDynamicSharedMemorySize, kR2Total));
      r2_configured_dev[dev_ & 63] = true;
    }
    const int work = (W_out / 256) * (H_out / 2) * p.n_tiles_n;
    const int pairs = work < g_max_sms / 2 ? work : g_max_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = kR2Total; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
    HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, k_conv_rows2, tmA, tmB, p));
    HIMO_LAUNCH_RET();
    return HIMO_OK;
  }
  // 256-channel layers on short rows: one 256-wide N tile per CTA pair (k_conv_wide)
  if (g_wide_tiles && P == 2 && d->Cout == 256 && d->Cin % kWideBK == 0 && !halo && m_tiles_total % 2 == 0 && (d->act == 0 || d->act == 1) &&
      !d->out_t && !d->mask_src && !d->stop_flag && d->b_group_k_stride == 0 && !d->b_k_total && !d->border_bias) {
    CUtensorMap tmBw, tmAw;
    cuuint64_t dimsw[3] = {(cuuint64_t)k_total, (cuuint64_t)d->Cout, 2};
    cuuint64_t stridesw[2] = {(cuuint64_t)k_total * 2, (cuuint64_t)d->Cout * k_total * 2};
    cuuint32_t boxw[3] = {(cuuint32_t)kWideBK, 128, 1};
    cuuint32_t estrw[3] = {1, 1, 1};
    if (enc(&tmBw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d->wgt, dimsw, stridesw, boxw, estrw,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return HIMO_ERR_ARG;
    {
      cuuint64_t dimsa[4] = {(cuuint64_t)d->Cin_total, (cuuint64_t)d->W_in, (cuuint64_t)d->H_in, (cuuint64_t)P};
      cuuint64_t stridesa[3] = {(cuuint64_t)d->Cin_total * 2, (cuuint64_t)d->W_in * d->Cin_total * 2,
                                (cuuint64_t)d->in_plane_stride * 2};
      cuuint32_t boxa[4] = {(cuuint32_t)kWideBK, (cuuint32_t)((TW - 1) * d->stride + 1), (cuuint32_t)((TH - 1) * d->stride + 1), 1};
      cuuint32_t estra[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
      if (enc(&tmAw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->in, dimsa, stridesa, boxa, estra,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return HIMO_ERR_ARG;
    }
    static bool wide_configured_dev[64] = {};
    int dev_ = 0;
    HIMO_CUDA_RET(cudaGetDevice(&dev_));
    bool& wide_configured = wide_configured_dev[dev_ & 63];
    if (!wide_configured) {
      HIMO_CUDA_RET(cudaFuncSetAttribute(k_conv_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideTotal));
      wide_configured = true;
    }
    const int work = (m_tiles_total / 2) * groups;
    const int pairs = work < kNumSMs / 2 ? work : kNumSMs / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = kWideTotal; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2 : 1;
    HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, k_conv_wide, tmAw, tmBw, p));
    HIMO_LAUNCH_RET();
    return HIMO_OK;
  }
  // weights-resident variants for the 64-channel encoder layers (one N tile, shared by all groups)
  if (g_weights_resident && BN == 64 && P == 2 && CGsel == 1 && d->Cout == 64 && d->b_group_k_stride == 0 && !d->b_k_total) {
    if (halo && taps * p.k_chunks == 18) return launch_conv<64, 2, 3, 1, 18>(tmA, tmB, p, stream);
    if (!halo && taps * p.k_chunks == 9) return launch_conv<64, 2, 1, 1, 9>(tmA, tmB, p, stream);
  }
  // streaming epilogue + TMA stores: split-plane 128-wide tiles whose whole K fits one accumulation chain (<= 96 hi*hi MMAs)
  {
    const int per_stage = 2 * (halo ? 3 : 1);
    const int k_iters = (taps / (halo ? 3 : 1)) * p.k_chunks;
    if (g_tma_store && !halo && P == 2 && BN == 128 && d->act < 5 && !d->out_fp32 && d->out_planes == 2 && !d->out_t &&
        !d->border_bias && d->out_group_pix_stride == 0 && !d->b_k_total && d->b_group_k_stride == 0 &&
        k_iters * per_stage <= 96 && d->Cout <= 1024 && d->cout_off % 8 == 0) {
      CUtensorMap tmO;
      cuuint64_t dimso[4] = {(cuuint64_t)d->Cout_total, (cuuint64_t)W_out, (cuuint64_t)H_out, 2};
      cuuint64_t strideso[3] = {(cuuint64_t)d->Cout_total * 2, (cuuint64_t)W_out * d->Cout_total * 2,
                                (cuuint64_t)d->out_plane_stride * 2};
      cuuint32_t boxo[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, 1};
      cuuint32_t estro[4] = {1, 1, 1, 1};
      if (enc(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d->out, dimso, strideso, boxo, estro, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return HIMO_ERR_ARG;
      p.flush_stages = k_iters;                  // one chain: the streaming epilogue reads the accumulators once
      if (CGsel == 2) return launch_conv<128, 2, 1, 2, 0, 1>(tmA, tmB, p, stream, &tmO);
      return launch_conv<128, 2, 1, 1, 0, 1>(tmA, tmB, p, stream, &tmO);
    }
  }
#define HIMO_CONV_CASE(bn, pp)                                                                        \
  if (BN == bn && P == pp) {                                                                          \
    if (CGsel == 2)                                                                                   \
      return halo ? launch_conv<bn, pp, 3, 2>(tmA, tmB, p, stream) : launch_conv<bn, pp, 1, 2>(tmA, tmB, p, stream); \
    return halo ? launch_conv<bn, pp, 3, 1>(tmA, tmB, p, stream) : launch_conv<bn, pp, 1, 1>(tmA, tmB, p, stream);   \
  }
  HIMO_CONV_CASE(128, 2)
  HIMO_CONV_CASE(96, 2)
  HIMO_CONV_CASE(64, 2)
  HIMO_CONV_CASE(128, 1)
  HIMO_CONV_CASE(96, 1)
  HIMO_CONV_CASE(64, 1)
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
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
  HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, k_upsample2x, (const __nv_bfloat16*)in, in_planes, in_plane_stride, h, w, c,
                                   (__nv_bfloat16*)out, out_planes, out_plane_stride, Cout_total, cout_off));
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}
