// himo_b200/csrc/decfused.cu -- H4 back end, fused: the whole per-point ConvGRU decoder in ONE persistent kernel.
//
// Replaces ConvGRUDecoder.forward_single + ConvGRU.forward (OSF/src/models/basic/decoder.py:177-237) after the
// gather: per GRU iteration  z,r = sigmoid(W_zr [h|x] + b),  q = tanh(W_q [r*h|x] + b),  h = (1-z) h + z q,
// then flow = W_2 GELU(W_0 [h|x] + b_0) + b_2.  The unfused path (csrc/dec.cu + five GEMM launches) moved every
// intermediate (z, r, q, r*h, h: ~1 GB per frame) through HBM and re-read the weights once per 128-point tile.
// Here a CTA pair (thread-block cluster of 2, tcgen05.mma.cta_group::2) owns 256 points for the whole decoder:
//   * the operand tile [h|x] lives in shared memory as split-fp16 planes in the K-major 64-byte-swizzled layout
//     tcgen05.mma reads; it is loaded once by TMA and then rewritten in place by the epilogue warps;
//   * accumulators live in tensor memory: R then Z in columns [0,192), Q in [192,384), the head GEMM in [384,448);
//   * r*h is produced 32 channels at a time into a two-slot ring that feeds the q GEMM while it is being computed;
//   * weights stream from L2 through a 3-stage TMA ring, each CTA fetching half of the rows (every weight byte
//     crosses the L2 fabric once per 256 points);
//   * warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warps 2-9 gate math (TMEM -> registers ->
//     shared memory) and the final 48->3 projection.
// Precision: split fp16 operands, three MMAs per k-step (hi*hi + hi*lo + lo*hi), fp32 accumulation; the hidden
// state between iterations is the 22-bit split value (relative 2^-22), well inside the 1e-4 flow budget.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "dec.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {

constexpr int kDfThreads = 320;
constexpr int kDfRows = 128;                       // points per CTA
constexpr int kDfTile = kDfRows * 64;              // one 32-channel plane tile: 8 KB
constexpr int kDfChunks = 9;                       // 288 / 32
constexpr int kDfHX = kDfChunks * 2 * kDfTile;     // 147456
constexpr int kDfRH = 2 * 2 * kDfTile;             // two ring slots x two planes
constexpr int kDfWRows = 96;                       // weight rows staged per CTA (N = 192 per pair)
constexpr int kDfWStage = 2 * kDfWRows * 64;       // 12288
constexpr int kDfWStages = 3;
constexpr int kDfOffRH = kDfHX;
constexpr int kDfOffW = kDfOffRH + kDfRH;
constexpr int kDfOffBar = kDfOffW + kDfWStages * kDfWStage;
constexpr int kDfOffConst = kDfOffBar + 512;
constexpr int kDfConstFloats = 192 * 3 + 64 + 3 * 48 + 4;
constexpr int kDfTotal = kDfOffConst + kDfConstFloats * 4 + 1024;   // + alignment slack
constexpr int kDfTmemCols = 512;
constexpr uint32_t kColA = 0, kColB = 192, kColC = 384;

struct DecFusedParams {
  int n;                 // valid rows (pc0 points)
  int n_pair_tiles;      // tiles of 256 rows
  int num_iters;
  const float* b_zr; const float* b_q; const float* b_0; const float* w2; const float* b2;
  float s_zr, s_q, s_0;
  const float4* pt4;
  float* flow;
};

// k order of every GEMM over [h|x]: the x chunks (always ready) first, then the h chunks in the order the
// state update of the previous iteration finishes them (the two column halves run in parallel)
__device__ __constant__ int kDfOrd[9] = {6, 7, 8, 0, 3, 1, 4, 2, 5};

__device__ __forceinline__ float df_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float df_tanh(float x) {
  const float e = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - e, 1.0f + e), x);
}
__device__ __forceinline__ float df_gelu(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// byte offset of 16-byte unit j (0..3) of `row` inside a [128 x 64 B] SWIZZLE_64B tile
__device__ __forceinline__ uint32_t df_swz(int row, int j) { return (uint32_t)(row * 64 + ((j ^ ((row >> 1) & 3)) << 4)); }

// 16 consecutive channels (units j0, j0+1) of one row: split planes -> fp32
__device__ __forceinline__ void df_load16(const uint8_t* tile_hi, const uint8_t* tile_lo, int row, int j0, float* v) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint4 a = *(const uint4*)(tile_hi + df_swz(row, j0 + u));
    const uint4 b = *(const uint4*)(tile_lo + df_swz(row, j0 + u));
    const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&aa[k]));
      const float2 l = __half22float2(*reinterpret_cast<const __half2*>(&bb[k]));
      v[u * 8 + 2 * k] = h.x + l.x;
      v[u * 8 + 2 * k + 1] = h.y + l.y;
    }
  }
}
__device__ __forceinline__ void df_store16(uint8_t* tile_hi, uint8_t* tile_lo, int row, int j0, const float* v) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::pack_split2(v[u * 8 + 2 * k], v[u * 8 + 2 * k + 1], true, hi[k], lo[k]);
    *(uint4*)(tile_hi + df_swz(row, j0 + u)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *(uint4*)(tile_lo + df_swz(row, j0 + u)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// The gate warps rendezvous on a named barrier and ONE thread signals the leader CTA's mbarrier: a cluster-scope
// release arrive costs a GPU-scope MEMBAR + ERRBAR (the profile showed every warp paying it per 32-channel chunk).
__device__ __forceinline__ void df_named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kDfThreads, 1)
k_dec_fused(const __grid_constant__ CUtensorMap tmHX, const __grid_constant__ CUtensorMap tmWzr,
            const __grid_constant__ CUtensorMap tmWq, const __grid_constant__ CUtensorMap tmW0,
            const DecFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sHX = smem;
  uint8_t* sRH = smem + kDfOffRH;
  uint8_t* sW = smem + kDfOffW;
  uint64_t* bars = (uint64_t*)(smem + kDfOffBar);
  uint64_t* w_full = bars;            // [3]
  uint64_t* w_empty = w_full + 3;     // [3]
  uint64_t* hx_tma = w_empty + 3;     // [9]
  uint64_t* hx_epi = hx_tma + 9;      // [6]
  uint64_t* r_full = hx_epi + 6;
  uint64_t* q_full = r_full + 1;
  uint64_t* z_full = q_full + 1;
  uint64_t* d_full = z_full + 1;
  uint64_t* rh_ready = d_full + 1;    // [2]
  uint64_t* rh_empty = rh_ready + 2;  // [2]
  uint64_t* r_empty = rh_empty + 2;
  uint64_t* zq_empty = r_empty + 1;
  uint64_t* d_empty = zq_empty + 1;
  uint64_t* tile_done = d_empty + 1;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tile_done + 1);
  float* cst = (float*)(smem + kDfOffConst);
  float* c_bz = cst;            // [192]
  float* c_br = cst + 192;      // [192]
  float* c_bq = cst + 384;      // [192]
  float* c_b0 = cst + 576;      // [64]
  float* c_w2 = cst + 640;      // [3][48]
  float* c_b2 = cst + 784;      // [3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = umma::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_workers = (int)(gridDim.x >> 1);
  const int worker = (int)(blockIdx.x >> 1);
  const int NI = p.num_iters;

  if (warp == 0 && lane == 0) {
    umma::tma_prefetch_desc(&tmHX); umma::tma_prefetch_desc(&tmWzr);
    umma::tma_prefetch_desc(&tmWq); umma::tma_prefetch_desc(&tmW0);
    for (int s = 0; s < 3; ++s) { umma::mbar_init(&w_full[s], 1); umma::mbar_init(&w_empty[s], 1); }
    for (int c = 0; c < 9; ++c) umma::mbar_init(&hx_tma[c], 1);
    for (int c = 0; c < 6; ++c) umma::mbar_init(&hx_epi[c], 2);          // one aggregated arrival per CTA
    umma::mbar_init(r_full, 1); umma::mbar_init(q_full, 1); umma::mbar_init(z_full, 1); umma::mbar_init(d_full, 1);
    for (int s = 0; s < 2; ++s) { umma::mbar_init(&rh_ready[s], 2); umma::mbar_init(&rh_empty[s], 1); }
    umma::mbar_init(r_empty, 2); umma::mbar_init(zq_empty, 2); umma::mbar_init(d_empty, 2);
    umma::mbar_init(tile_done, 1);
    umma::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 192; i += kDfThreads) {
    c_bz[i] = __ldg(p.b_zr + i); c_br[i] = __ldg(p.b_zr + 192 + i); c_bq[i] = __ldg(p.b_q + i);
  }
  for (int i = threadIdx.x; i < 64; i += kDfThreads) c_b0[i] = __ldg(p.b_0 + i);
  for (int i = threadIdx.x; i < 144; i += kDfThreads) c_w2[i] = __ldg(p.w2 + i);
  if (threadIdx.x < 3) c_b2[threadIdx.x] = __ldg(p.b2 + threadIdx.x);
  __syncthreads();
  umma::cluster_sync();
  // tcgen05.alloc.cta_group::2 is a compiler-generated handshake through the PEER CTA's reserved shared memory (remote
  // mbarrier arrive + remote store of the address): it may only run once the peer CTA is known to be executing, i.e. after a
  // cluster barrier.  Allocating before it hangs when the two CTAs of a pair start far apart, which several streams' kernels
  // sharing the GPU provoke (profiles/r02_two_cta_alloc_hang.txt).
  if (warp == 1) umma::tmem_alloc_2cta(tmem_ptr_smem, kDfTmemCols);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t git = 0;
    auto stream_w = [&](const CUtensorMap* tm, int row_base, int rows, int c) {
      const int s = git % kDfWStages;
      const uint32_t ph = (git / kDfWStages) & 1;
      umma::mbar_wait(&w_empty[s], ph ^ 1);
      const uint32_t fb = umma::mapa_u32(umma::smem_u32(&w_full[s]), 0);
      uint8_t* dst = sW + s * kDfWStage;
      if (umma::elect_one()) {
        if (leader) umma::mbar_arrive_expect_tx(&w_full[s], (uint32_t)(2 * rows * 64 * 2));
        umma::tma_load_3d_2cta(dst, tm, fb, c * 32, row_base, 0);
        umma::tma_load_3d_2cta(dst + rows * 64, tm, fb, c * 32, row_base, 1);
      }
      __syncwarp();
      ++git;
    };
    int tc = 0;
    for (int tile = worker; tile < p.n_pair_tiles; tile += n_workers, ++tc) {
      if (tc > 0) umma::mbar_wait(tile_done, (uint32_t)((tc - 1) & 1));   // the previous tile's MMAs have read [h|x]
      const int row0 = (tile * 2 + (int)cta_rank) * kDfRows;
      for (int i = 0; i < 9; ++i) {
        const int c = kDfOrd[i];
        const uint32_t fb = umma::mapa_u32(umma::smem_u32(&hx_tma[c]), 0);
        if (umma::elect_one()) {
          if (leader) umma::mbar_arrive_expect_tx(&hx_tma[c], (uint32_t)(2 * 2 * kDfTile));
          umma::tma_load_3d_2cta(sHX + (c * 2) * kDfTile, &tmHX, fb, c * 32, row0, 0);
          umma::tma_load_3d_2cta(sHX + (c * 2 + 1) * kDfTile, &tmHX, fb, c * 32, row0, 1);
        }
        __syncwarp();
      }
      for (int it = 0; it < NI; ++it) {
        for (int i = 0; i < 9; ++i) stream_w(&tmWzr, 192 + (int)cta_rank * kDfWRows, kDfWRows, kDfOrd[i]);   // r rows
        for (int c = 6; c < 9; ++c) stream_w(&tmWq, (int)cta_rank * kDfWRows, kDfWRows, c);
        for (int c = 0; c < 6; ++c) stream_w(&tmWq, (int)cta_rank * kDfWRows, kDfWRows, c);
        for (int i = 0; i < 9; ++i) stream_w(&tmWzr, (int)cta_rank * kDfWRows, kDfWRows, kDfOrd[i]);         // z rows
      }
      for (int i = 0; i < 9; ++i) stream_w(&tmW0, (int)cta_rank * 32, 32, kDfOrd[i]);
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc192 = umma::idesc_f16kind_f32(256, 192, 0u, 0u);
      constexpr uint32_t idesc64 = umma::idesc_f16kind_f32(256, 64, 0u, 0u);
      uint32_t git = 0, gc = 0;
      // one k-chunk (32 channels): cross terms first, then hi*hi, all into the same accumulator
      auto chunk_mma = [&](uint32_t a_hi, uint32_t a_lo, int rows, uint32_t tmem_d, uint32_t idesc, bool first) {
        const int s = git % kDfWStages;
        const uint32_t ph = (git / kDfWStages) & 1;
        umma::mbar_wait(&w_full[s], ph);
        umma::tc_fence_after();
        const uint32_t b_hi = umma::smem_u32(sW + s * kDfWStage), b_lo = b_hi + rows * 64;
        if (umma::elect_one()) {
          const uint64_t dah = umma::smem_desc_kmajor<64>(a_hi), dal = umma::smem_desc_kmajor<64>(a_lo);
          const uint64_t dbh = umma::smem_desc_kmajor<64>(b_hi), dbl = umma::smem_desc_kmajor<64>(b_lo);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t koff = (uint64_t)(k * 32 >> 4);
            umma::mma_bf16_ss_2cta(tmem_d, dah + koff, dbl + koff, idesc, (first && k == 0) ? 0u : 1u);
            umma::mma_bf16_ss_2cta(tmem_d, dal + koff, dbh + koff, idesc, 1u);
            umma::mma_bf16_ss_2cta(tmem_d, dah + koff, dbh + koff, idesc, 1u);
          }
          umma::mma_commit_2cta(&w_empty[s]);
        }
        __syncwarp();
        ++git;
      };
      auto commit = [&](uint64_t* bar) {
        if (umma::elect_one()) umma::mma_commit_2cta(bar);
        __syncwarp();
      };
      auto hx_chunk_ready = [&](int c, int tc, int n_prev) {   // n_prev < 0: the TMA load of this tile
        if (c >= 6 || n_prev < 0) umma::mbar_wait(&hx_tma[c], (uint32_t)(tc & 1));
        else umma::mbar_wait(&hx_epi[c], (uint32_t)(n_prev & 1));
      };
      const uint32_t hx_addr = umma::smem_u32(sHX), rh_addr = umma::smem_u32(sRH);
      int tc = 0;
      for (int tile = worker; tile < p.n_pair_tiles; tile += n_workers, ++tc) {
        for (int it = 0; it < NI; ++it) {
          const int n_it = tc * NI + it;
          umma::mbar_wait(zq_empty, (uint32_t)((n_it & 1) ^ 1));          // the previous update has read Z and Q
          umma::tc_fence_after();
          // ---- R = W_r [h|x]
          for (int i = 0; i < 9; ++i) {
            const int c = kDfOrd[i];
            hx_chunk_ready(c, tc, it == 0 ? -1 : n_it - 1);
            chunk_mma(hx_addr + (c * 2) * kDfTile, hx_addr + (c * 2 + 1) * kDfTile, kDfWRows, tmem_base + kColA, idesc192, i == 0);
          }
          commit(r_full);
          // ---- Q = W_q [r*h|x]: the x part first, then the r*h chunks as the gate warps produce them
          for (int c = 6; c < 9; ++c)
            chunk_mma(hx_addr + (c * 2) * kDfTile, hx_addr + (c * 2 + 1) * kDfTile, kDfWRows, tmem_base + kColB, idesc192, c == 6);
          for (int c = 0; c < 6; ++c, ++gc) {
            const int slot = gc & 1;
            umma::mbar_wait(&rh_ready[slot], (gc >> 1) & 1);
            chunk_mma(rh_addr + (slot * 2) * kDfTile, rh_addr + (slot * 2 + 1) * kDfTile, kDfWRows, tmem_base + kColB, idesc192, false);
            commit(&rh_empty[slot]);
          }
          commit(q_full);
          // ---- Z = W_z [h|x] into R's columns once the gate warps have consumed R
          umma::mbar_wait(r_empty, (uint32_t)(n_it & 1));
          umma::tc_fence_after();
          for (int i = 0; i < 9; ++i) {
            const int c = kDfOrd[i];
            chunk_mma(hx_addr + (c * 2) * kDfTile, hx_addr + (c * 2 + 1) * kDfTile, kDfWRows, tmem_base + kColA, idesc192, i == 0);
          }
          commit(z_full);
        }
        // ---- head: Y = W_0 [h|x]
        umma::mbar_wait(d_empty, (uint32_t)((tc & 1) ^ 1));
        umma::tc_fence_after();
        for (int i = 0; i < 9; ++i) {
          const int c = kDfOrd[i];
          hx_chunk_ready(c, tc, NI > 0 ? tc * NI + NI - 1 : -1);
          chunk_mma(hx_addr + (c * 2) * kDfTile, hx_addr + (c * 2 + 1) * kDfTile, 32, tmem_base + kColC, idesc64, i == 0);
        }
        commit(d_full);
        commit(tile_done);
      }
    }
  } else {
    // ===================== gate / update / head warps =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_rh_ready0 = umma::mapa_u32(umma::smem_u32(&rh_ready[0]), 0);
    const uint32_t a_rh_ready1 = umma::mapa_u32(umma::smem_u32(&rh_ready[1]), 0);
    const uint32_t a_r_empty = umma::mapa_u32(umma::smem_u32(r_empty), 0);
    const uint32_t a_zq_empty = umma::mapa_u32(umma::smem_u32(zq_empty), 0);
    const uint32_t a_d_empty = umma::mapa_u32(umma::smem_u32(d_empty), 0);
    const uint32_t a_hx_epi0 = umma::mapa_u32(umma::smem_u32(&hx_epi[0]), 0);   // barriers are 8 bytes apart
    uint32_t gc = 0;
    int tc = 0;
    for (int tile = worker; tile < p.n_pair_tiles; tile += n_workers, ++tc) {
      for (int it = 0; it < NI; ++it) {
        const int n_it = tc * NI + it;
        // ---- r = sigmoid(R), r*h -> ring
        umma::mbar_wait(r_full, (uint32_t)(n_it & 1));
        umma::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 6; ++c, ++gc) {
          const int slot = gc & 1;
          uint32_t acc[16];
          umma::tmem_ld_32x16(tlane + kColA + (uint32_t)(c * 32 + half * 16), acc);
          float h[16];
          df_load16(sHX + (c * 2) * kDfTile, sHX + (c * 2 + 1) * kDfTile, row, half * 2, h);
          umma::tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = df_sigmoid(__fmaf_rn(__uint_as_float(acc[j]), p.s_zr, c_br[c * 32 + half * 16 + j])) * h[j];
          umma::mbar_wait(&rh_empty[slot], ((gc >> 1) & 1) ^ 1);          // the q GEMM has read this slot's previous chunk
          df_store16(sRH + (slot * 2) * kDfTile, sRH + (slot * 2 + 1) * kDfTile, row, half * 2, v);
          umma::fence_proxy_async();
          df_named_sync(1, 256);
          if (warp == 2 && lane == 0) umma::mbar_arrive_cluster(slot ? a_rh_ready1 : a_rh_ready0);
        }
        umma::tc_fence_before();
        df_named_sync(1, 256);
        if (warp == 2 && lane == 0) umma::mbar_arrive_cluster(a_r_empty);
        // ---- z = sigmoid(Z), q = tanh(Q), h <- (1-z) h + z q, written back into the operand tile
        umma::mbar_wait(z_full, (uint32_t)(n_it & 1));                     // (the commit also covers the q GEMM)
        umma::tc_fence_after();
#pragma unroll 1
        for (int g = 0; g < 6; ++g) {
          const int col = half * 96 + g * 16;
          const int c = col >> 5, j0 = ((col & 31) >> 4) * 2;
          uint32_t az[16], aq[16];
          umma::tmem_ld_32x16(tlane + kColA + (uint32_t)col, az);
          umma::tmem_ld_32x16(tlane + kColB + (uint32_t)col, aq);
          float h[16];
          df_load16(sHX + (c * 2) * kDfTile, sHX + (c * 2 + 1) * kDfTile, row, j0, h);
          umma::tmem_ld_wait();
          float hn[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float z = df_sigmoid(__fmaf_rn(__uint_as_float(az[j]), p.s_zr, c_bz[col + j]));
            const float qq = df_tanh(__fmaf_rn(__uint_as_float(aq[j]), p.s_q, c_bq[col + j]));
            hn[j] = __fadd_rn(__fmul_rn(1.0f - z, h[j]), __fmul_rn(z, qq));
          }
          df_store16(sHX + (c * 2) * kDfTile, sHX + (c * 2 + 1) * kDfTile, row, j0, hn);
          if (g & 1) {   // both 16-column groups of chunk c are written: hand it to the next GEMM
            umma::fence_proxy_async();
            df_named_sync(2 + half, 128);                          // the four warps that own this column half
            if ((warp == 2 || warp == 6) && lane == 0) umma::mbar_arrive_cluster(a_hx_epi0 + 8u * (uint32_t)c);
          }
        }
        umma::tc_fence_before();
        df_named_sync(1, 256);
        if (warp == 2 && lane == 0) umma::mbar_arrive_cluster(a_zq_empty);
      }
      // ---- head: flow = W_2 GELU(Y + b_0) + b_2 (one warp per TMEM lane quadrant)
      if (half == 0) {
        umma::mbar_wait(d_full, (uint32_t)(tc & 1));
        umma::tc_fence_after();
        float o0 = c_b2[0], o1 = c_b2[1], o2 = c_b2[2];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          uint32_t acc[16];
          umma::tmem_ld_32x16(tlane + kColC + (uint32_t)(g * 16), acc);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = g * 16 + j;
            const float y = df_gelu(__fmaf_rn(__uint_as_float(acc[j]), p.s_0, c_b0[k]));
            o0 = fmaf(c_w2[k], y, o0); o1 = fmaf(c_w2[48 + k], y, o1); o2 = fmaf(c_w2[96 + k], y, o2);
          }
        }
        const long long gi = (long long)(tile * 2 + (int)cta_rank) * kDfRows + row;
        if (gi < p.n) {
          const bool valid = __float_as_int(p.pt4[gi].w) >= 0;
          p.flow[3 * gi] = valid ? o0 : 0.f; p.flow[3 * gi + 1] = valid ? o1 : 0.f; p.flow[3 * gi + 2] = valid ? o2 : 0.f;
        }
        umma::tc_fence_before();
        df_named_sync(2, 128);
        if (warp == 2 && lane == 0) umma::mbar_arrive_cluster(a_d_empty);
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync();
  if (warp == 1) umma::tmem_dealloc_2cta(tmem_base, kDfTmemCols);
}

static PFN_cuTensorMapEncodeTiled df_encode_fn() {
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

static bool df_map(CUtensorMap* tm, const void* base, long long k, long long rows, long long plane_stride_elems,
                   int box_rows) {
  PFN_cuTensorMapEncodeTiled enc = df_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)k * 2, (cuuint64_t)plane_stride_elems * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// hx: [2][n_pad][288] split planes (n_pad a multiple of 256); weights as in himo_deflowpp_weights (planes == 2).
int dec_fused(const __nv_bfloat16* hx, int n, int n_pad, int num_iters, const himo_deflowpp_weights* w,
              const float4* pt4, float* flow, cudaStream_t stream) {
  if (n <= 0) return HIMO_OK;
  if (n_pad % 256 || n_pad < n || w->planes != 2 || num_iters < 0) return HIMO_ERR_ARG;
  CUtensorMap tmHX, tmWzr, tmWq, tmW0;
  if (!df_map(&tmHX, hx, 288, n_pad, (long long)n_pad * 288, kDfRows) ||
      !df_map(&tmWzr, w->gru_zr_w, 288, 384, 384ll * 288, kDfWRows) ||
      !df_map(&tmWq, w->gru_q_w, 288, 192, 192ll * 288, kDfWRows) ||
      !df_map(&tmW0, w->dec0_w, 288, 64, 64ll * 288, 32))
    return HIMO_ERR_UNSUPPORTED;
  DecFusedParams p;
  p.n = n; p.n_pair_tiles = n_pad / 256; p.num_iters = num_iters;
  p.b_zr = w->gru_zr_b; p.b_q = w->gru_q_b; p.b_0 = w->dec0_b; p.w2 = w->dec2_w; p.b2 = w->dec2_b;
  p.s_zr = w->gru_zr_s != 0.f ? w->gru_zr_s : 1.f;
  p.s_q = w->gru_q_s != 0.f ? w->gru_q_s : 1.f;
  p.s_0 = w->dec0_s != 0.f ? w->dec0_s : 1.f;
  p.pt4 = pt4; p.flow = flow;
  static bool configured_dev[64] = {};      // the attribute is per device: one flag per device ordinal
  int dev_ = 0;
  HIMO_CUDA_RET(cudaGetDevice(&dev_));
  bool& configured = configured_dev[dev_ & 63];
  if (!configured) {
    HIMO_CUDA_RET(cudaFuncSetAttribute(k_dec_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, kDfTotal));
    configured = true;
  }
  const int pairs = p.n_pair_tiles < kNumSMs / 2 ? p.n_pair_tiles : kNumSMs / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(kDfThreads); cfg.dynamicSmemBytes = kDfTotal; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, k_dec_fused, tmHX, tmWzr, tmWq, tmW0, p));
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

}  // namespace himo
