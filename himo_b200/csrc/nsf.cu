// himo_b200/csrc/nsf.cu -- H3: FastNSF per-frame-pair optimisation on the device.
//
// Replaces FastNSF.optimize (OSF/src/models/fastnsf.py:105-169) with its pieces
//   DT.__init__ / FastGeodis.generalised_geodesic3d   fastnsf.py:30-57  (distance volume)
//   DT.torch_bilinear_distance (+ autograd)            fastnsf.py:59-80  (trilinear lookup)
//   Neural_Prior.forward (+ autograd)                  basic/nsfp_module.py:41-47 (3->128x8->3 ReLU MLP)
//   torch.optim.Adam.step, EarlyStopping.step          fastnsf.py:134-164, nsfp_module.py:65-82
// The reference runs ~60 launches and >= 2 host syncs per iteration.  Here one iteration is 34 launches
// with no host involvement: the seven 128x128 hidden layers run forward, backward (dX) and weight-gradient
// (dW, split-K) on the tcgen05 GEMM kernel of conv.cu with ReLU / ReLU-mask / transposed stores fused into
// its epilogue; loss, best-flow snapshot, early stopping and Adam live in a device control block; the
// host only polls a stop flag every few iterations.
#include <cstdlib>
#include <cudaTypedefs.h>
#include <math.h>

#include "common.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {

constexpr int kNsfHidden = 128;
constexpr int kNsfLayers = 8;                 // hidden layers (nn_layers.0 .. nn_layers.14)
constexpr float kNsfWScale = 256.f;           // power-of-two prescale of the packed weights (split mode)

struct NsfCtl {
  int stop;            // set by the early-stopping state machine
  int iters;           // iterations executed (loss evaluations)
  int snapshot;        // this iteration's flow is the new best
  int es_has_best;
  int es_bad;
  float best_loss;
  float es_best;
  float loss;
};

struct NsfVol {        // distance volume geometry (host + device copy)
  float lo[3];
  int dims[3];         // H, W, D  (x, y, z)
  float gf;
};

// ------------------------------------------------------------------------------ DT: bounds
__device__ __forceinline__ unsigned nsf_f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float nsf_ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void k_nsf_bbox_init(unsigned* b) {
  if (threadIdx.x < 3) b[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) b[threadIdx.x] = 0u;
}
__global__ void __launch_bounds__(256)
k_nsf_bbox(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, unsigned* __restrict__ bbox) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += gridDim.x * blockDim.x) {
    const float* p = i < na ? a + 3 * (size_t)i : b + 3 * (size_t)(i - na);
#pragma unroll
    for (int k = 0; k < 3; ++k) { float v = __ldg(p + k); lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (lo[k] <= hi[k]) { atomicMin(bbox + k, nsf_f2ord(lo[k])); atomicMax(bbox + 3 + k, nsf_f2ord(hi[k])); }
}
// lo = floor(min*gf - 1)/gf, hi = ceil(max*gf + 1)/gf, samples = ceil((hi-lo)*gf) + 2, all in fp32
// (fastnsf.py:120-126 and DT.__init__ :34-36).
__global__ void k_nsf_vol(const unsigned* __restrict__ bbox, float gf, NsfVol* __restrict__ v) {
  if (threadIdx.x != 0) return;
  for (int k = 0; k < 3; ++k) {
    const float mn = nsf_ord2f(bbox[k]), mx = nsf_ord2f(bbox[3 + k]);
    const float lo = __fdiv_rn(floorf(__fsub_rn(__fmul_rn(mn, gf), 1.f)), gf);
    const float hi = __fdiv_rn(ceilf(__fadd_rn(__fmul_rn(mx, gf), 1.f)), gf);
    v->lo[k] = lo;
    v->dims[k] = (int)ceilf(__fmul_rn(__fsub_rn(hi, lo), gf)) + 2;
  }
  v->gf = gf;
}

// ------------------------------------------------------------------------------ DT: occupancy + raster passes
__global__ void __launch_bounds__(256)
k_nsf_fill(float* __restrict__ d, long long n, float v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = v;
}
__global__ void __launch_bounds__(256)
k_nsf_occupancy(const float* __restrict__ pc1, int n, NsfVol v, float* __restrict__ d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // index = round((p - V[0]) * gf), torch.round = half-to-even = cvt.rni (fastnsf.py:46-50)
    const int ix = __float2int_rn(__fmul_rn(__fsub_rn(pc1[3 * (size_t)i], v.lo[0]), v.gf));
    const int iy = __float2int_rn(__fmul_rn(__fsub_rn(pc1[3 * (size_t)i + 1], v.lo[1]), v.gf));
    const int iz = __float2int_rn(__fmul_rn(__fsub_rn(pc1[3 * (size_t)i + 2], v.lo[2]), v.gf));
    if (ix >= 0 && ix < v.dims[0] && iy >= 0 && iy < v.dims[1] && iz >= 0 && iz < v.dims[2])
      d[((size_t)ix * v.dims[1] + iy) * v.dims[2] + iz] = 0.f;
  }
}

// One raster pass of the FastGeodis Euclidean transform advances along `axis`; plane p takes
//   new[p] = min(old[p], min over the 3x3 neighbourhood of new[p -/+ 1] + step length).
// A block owns a kDtTile x kDtTile tile of the plane and advances kDtSteps planes per launch from a halo of kDtSteps
// cells (the dependency cone widens by one cell per plane), so a pass costs n_axis/kDtSteps launches
// instead of n_axis.  Halo cells may read a neighbour block's already-updated value of plane p; because
// the update is an idempotent min over the same inputs this cannot change the result.
constexpr int kDtStepsDefault = 16;
constexpr int kDtTile = 16;   // small tiles: the axis-0/1 planes are only ~50 k cells, 32x32 tiles left most SMs idle
constexpr int kDtTileBig = 32;   // optional for planes of >= 256 k cells (himo_nsf_set_dt_big_tiles): 4x instead of 9x recomputation, but slower

template <int kDtTile, int kDtSteps>
__global__ void __launch_bounds__(256)
k_nsf_dt_pass(float* __restrict__ d, int n0, int n1, int n2, int axis, int dir, int p_begin, int count,
              float l00, float l01, float l11) {
  // l00 = straight step, l01 = one lateral move, l11 = two lateral moves
  constexpr int kDtReg = kDtTile + 2 * kDtSteps;   // 48 / 64
  __shared__ float buf[2][kDtReg][kDtReg + 1];
  const int n[3] = {n0, n1, n2};
  const long long st[3] = {(long long)n1 * n2, (long long)n2, 1};
  // in-plane axes: w (the fast thread index) is always the higher-numbered axis = the smaller memory stride, so
  // that a warp reads consecutive addresses (with (axis+1)%3, (axis+2)%3 the axis-1 passes strode n1*n2 floats)
  const int a = axis, h = axis == 0 ? 1 : 0, w = axis == 2 ? 1 : 2;
  const int h0 = blockIdx.y * kDtTile - kDtSteps, w0 = blockIdx.x * kDtTile - kDtSteps;
  const int pp = p_begin - dir;     // plane already final
  for (int t = threadIdx.x; t < kDtReg * kDtReg; t += blockDim.x) {
    const int rh = t / kDtReg, rw = t % kDtReg;
    const int ih = h0 + rh, iw = w0 + rw;
    float v = INFINITY;
    if (ih >= 0 && ih < n[h] && iw >= 0 && iw < n[w]) v = d[pp * st[a] + ih * st[h] + iw * st[w]];
    buf[0][rh][rw] = v;
  }
  __syncthreads();
  int cur = 0;
  for (int s = 0; s < count; ++s) {
    const int p = p_begin + s * dir;
    const int shrink = s + 1;                       // cells still exact after s+1 steps
    const int lo = shrink, hi = kDtReg - shrink;     // region [lo,hi) in both directions
    const int side = hi - lo;
    for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
      const int rh = lo + t / side, rw = lo + t % side;
      const int ih = h0 + rh, iw = w0 + rw;
      float best = INFINITY;
      if (ih >= 0 && ih < n[h] && iw >= 0 && iw < n[w]) {
        best = d[p * st[a] + ih * st[h] + iw * st[w]];
#pragma unroll
        for (int dh = -1; dh <= 1; ++dh)
#pragma unroll
          for (int dw = -1; dw <= 1; ++dw) {
            const float step = (dh == 0 && dw == 0) ? l00 : ((dh != 0 && dw != 0) ? l11 : l01);
            best = fminf(best, __fadd_rn(buf[cur][rh + dh][rw + dw], step));
          }
        if (rh >= kDtSteps && rh < kDtSteps + kDtTile && rw >= kDtSteps && rw < kDtSteps + kDtTile)
          d[p * st[a] + ih * st[h] + iw * st[w]] = best;
      }
      buf[cur ^ 1][rh][rw] = best;
    }
    __syncthreads();
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------ DT: one cluster sweeps a whole pass
// k_nsf_dt_sweep: the same recurrence as k_nsf_dt_pass for the two passes whose planes are small (axis 0 and 1: ~53 k cells
// of the 1040 x 1030 x 52 volume), as ONE launch per (axis, direction) instead of 65.  A thread-block cluster owns the
// plane: CTA r keeps rows [r R, (r+1) R) of the last three planes in shared memory, computes its rows of plane p from the
// previous plane and sends its first / last row into the neighbours' halo rows with st.async (distributed shared memory),
// which also signals the receiver's mbarrier for that plane.  There is no cluster barrier and no memory fence in the loop:
// a step waits for its two halo rows (mbarrier) and for its own rows (__syncthreads).  Three state buffers make that safe: the
// neighbour can run at most one step ahead (it needs my row of step s to start step s+1), and what it then writes is
// generation s+2, a different buffer from the generation s that my slower warps may still be reading.
// The old values of the next kDtcPF planes are prefetched with cp.async into a per-thread-private ring.
// min / add with the same three constants in a different association: fminf is exact and x -> fl(x + c) is monotonic,
// so min(fl(a+c), fl(b+c)) == fl(min(a,b) + c) and the result is bit-identical to k_nsf_dt_pass.
constexpr int kDtcThreads = 1024;
constexpr int kDtcPF = 4;
constexpr int kDtcVS = 4;          // a thread owns a vertical strip of 4 cells: 18 shared-memory loads instead of 36
constexpr int kDtcMaxStrips = 2;

__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cp_async_f32(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(umma::smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kDtcThreads, 1)
k_nsf_dt_sweep(float* __restrict__ d, int n0, int n1, int n2, int axis, int dir, float l00, float l01, float l11, int R,
               int csize, int SW, long long* __restrict__ dbg) {
  extern __shared__ float dtc_smem[];
  __shared__ uint64_t halo_bar[3];
  const int n[3] = {n0, n1, n2};
  const long long stp = axis == 0 ? (long long)n1 * n2 : (long long)n2;       // stride of the sweep axis
  const int sth = axis == 0 ? n2 : n1 * n2;                                   // stride of the in-plane row axis (y or x)
  const int nh = axis == 0 ? n1 : n0, nw = n2, np = n[axis];
  const int rank = (int)umma::cluster_ctarank();
  const int row0 = rank * R;
  // (written as comparisons of the inputs: nvcc 12.9 folds `max(0, min(R, x)) == R` into the predicate output of
  //  VIMNMX.RELU and gets it wrong -- true for x <= R -- found with a device printf, profiles/r02_dt_cluster_sweep.txt)
  const int left = nh - row0;
  const int rows = left <= 0 ? 0 : (left < R ? left : R);
  const bool talk_up = rank > 0 && left > 0;                   // rank-1 (always full then) and I exchange rows
  const bool talk_down = rank + 1 < csize && left - R > 0;     // rank+1 has rows (then I am full)
  const int SB = (R + 2) * SW;                                 // SW: state row stride (>= nw + 2, dt_sweep_row_stride), state buffer size
  float* ring = dtc_smem + 3 * (size_t)SB;                     // [kDtcPF][R * nw]
  const int tid = threadIdx.x;
  const int ngrp = (R + kDtcVS - 1) / kDtcVS;
  int sbase[kDtcMaxStrips], gbase[kDtcMaxStrips], cbase[kDtcMaxStrips], nrow[kDtcMaxStrips];
  uint32_t send_up = 0, send_dn = 0;                           // bit u*kDtcVS+j: that cell sits in my first / last row
  int col1[kDtcMaxStrips];
#pragma unroll
  for (int u = 0; u < kDtcMaxStrips; ++u) {
    const int strip = tid + u * kDtcThreads;
    const int g = strip / nw, col = strip - g * nw;
    const int rb = g * kDtcVS;
    sbase[u] = (rb + 1) * SW + col + 1;
    gbase[u] = (row0 + rb) * sth + col;
    cbase[u] = rb * nw + col;
    const int nr = g < ngrp ? rows - rb : 0;
    nrow[u] = nr <= 0 ? 0 : (nr < kDtcVS ? nr : kDtcVS);
    col1[u] = col + 1;
    if (nrow[u] > 0 && rb == 0 && talk_up) send_up |= 1u << (u * kDtcVS);
    if (nrow[u] > 0 && talk_down && rb <= R - 1 && R - 1 < rb + kDtcVS) send_dn |= 1u << (u * kDtcVS + R - 1 - rb);
  }
  for (int i = tid; i < 3 * SB; i += kDtcThreads) dtc_smem[i] = INFINITY;
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) umma::mbar_init(&halo_bar[b], 1);
    umma::fence_barrier_init();
  }
  __syncthreads();
  umma::cluster_sync();                 // every CTA runs, has cleared its buffers and initialised its barriers
  const uint32_t s_base = umma::smem_u32(dtc_smem), bar_base = umma::smem_u32(&halo_bar[0]);
  const uint32_t up_s = talk_up ? umma::mapa_u32(s_base, (uint32_t)(rank - 1)) : 0u;
  const uint32_t up_b = talk_up ? umma::mapa_u32(bar_base, (uint32_t)(rank - 1)) : 0u;
  const uint32_t dn_s = talk_down ? umma::mapa_u32(s_base, (uint32_t)(rank + 1)) : 0u;
  const uint32_t dn_b = talk_down ? umma::mapa_u32(bar_base, (uint32_t)(rank + 1)) : 0u;
  const uint32_t halo_bytes = (uint32_t)(((talk_up ? 1 : 0) + (talk_down ? 1 : 0)) * nw * 4);
  // my first row -> row R+1 of rank-1, my last row -> row 0 of rank+1, in generation buffer `buf`
  auto send = [&](int buf, int u, int j, float v) {
    const uint32_t bit = 1u << (u * kDtcVS + j);
    if (send_up & bit) st_async_f32(up_s + (uint32_t)((buf * SB + (R + 1) * SW + col1[u]) * 4), v, up_b + 8u * buf);
    if (send_dn & bit) st_async_f32(dn_s + (uint32_t)((buf * SB + col1[u]) * 4), v, dn_b + 8u * buf);
  };
  const int count = np - 1;
  const int p_begin = dir > 0 ? 1 : np - 2;
  if (tid == 0) umma::mbar_arrive_expect_tx(&halo_bar[0], halo_bytes);
  {
    const float* src = d + (long long)(p_begin - dir) * stp;         // the plane that is already final: generation 0
#pragma unroll
    for (int u = 0; u < kDtcMaxStrips; ++u)
#pragma unroll
      for (int j = 0; j < kDtcVS; ++j)
        if (j < nrow[u]) {
          const float v = src[gbase[u] + j * sth];
          dtc_smem[sbase[u] + j * SW] = v;
          send(0, u, j, v);
        }
  }
  auto prefetch = [&](int slot, int plane) {
    const float* src = d + (long long)plane * stp;
    float* dst = ring + (size_t)slot * R * nw;
#pragma unroll
    for (int u = 0; u < kDtcMaxStrips; ++u) {
      if (nrow[u] == kDtcVS) {
        float* r4 = dst + cbase[u];
        const float* g4 = src + gbase[u];
#pragma unroll
        for (int j = 0; j < kDtcVS; ++j) cp_async_f32(r4 + j * nw, g4 + j * sth);
      } else {
#pragma unroll
        for (int j = 0; j < kDtcVS; ++j)
          if (j < nrow[u]) cp_async_f32(dst + cbase[u] + j * nw, src + gbase[u] + j * sth);
      }
    }
  };
#pragma unroll
  for (int j = 0; j < kDtcPF; ++j) {
    if (j < count) prefetch(j, p_begin + j * dir);
    cp_async_commit();
  }
  __syncthreads();
  int cur = 0;                                    // buffer of generation s
  long long t_cp = 0, t_halo = 0, t_comp = 0, t_sync = 0, t0 = 0;
  for (int s = 0; s < count; ++s) {
    const int p = p_begin + s * dir;
    const int slot = s % kDtcPF;
    const int nxt = cur == 2 ? 0 : cur + 1;
    if (tid == 0 && s + 1 < count) umma::mbar_arrive_expect_tx(&halo_bar[nxt], halo_bytes);
    if (dbg) t0 = clock64();
    cp_async_wait<kDtcPF - 1>();
    if (dbg) { const long long t = clock64(); t_cp += t - t0; t0 = t; }
    umma::mbar_wait(&halo_bar[cur], (uint32_t)((s / 3) & 1));        // the neighbours' rows of generation s are here
    if (dbg) { const long long t = clock64(); t_halo += t - t0; t0 = t; }
    const float* prev = dtc_smem + cur * SB;
    float* out = dtc_smem + nxt * SB;
    float* dst = d + (long long)p * stp;
    const float* old = ring + (size_t)slot * R * nw;
    const bool talk = s + 1 < count;              // nobody reads the halos of the last generation
#pragma unroll
    for (int u = 0; u < kDtcMaxStrips; ++u) {
      if (nrow[u] == kDtcVS) {                      // full strip: straight-line code, no per-cell predicates
        const float* q = prev + sbase[u] - SW;      // row above the strip
        float a[kDtcVS + 2], m[kDtcVS + 2];
#pragma unroll
        for (int i = 0; i < kDtcVS + 2; ++i) {
          a[i] = q[i * SW];
          m[i] = fminf(q[i * SW - 1], q[i * SW + 1]);
        }
        const float* o4 = old + cbase[u];
        float* w4 = out + sbase[u];
        float* g4 = dst + gbase[u];
        float best[kDtcVS];
#pragma unroll
        for (int j = 0; j < kDtcVS; ++j) {
          const float e = fminf(m[j + 1], fminf(a[j], a[j + 2]));
          const float x = fminf(m[j], m[j + 2]);
          best[j] = fminf(fminf(o4[j * nw], __fadd_rn(a[j + 1], l00)), fminf(__fadd_rn(e, l01), __fadd_rn(x, l11)));
          w4[j * SW] = best[j];
          g4[j * sth] = best[j];
        }
        if (talk && ((send_up | send_dn) >> (u * kDtcVS) & 0xfu)) {
#pragma unroll
          for (int j = 0; j < kDtcVS; ++j) send(nxt, u, j, best[j]);
        }
      } else if (nrow[u] > 0) {                     // ragged last strip of a CTA
        const float* q = prev + sbase[u] - SW;
        float a[kDtcVS + 2], m[kDtcVS + 2];
#pragma unroll
        for (int i = 0; i < kDtcVS + 2; ++i) {
          if (i <= nrow[u] + 1) {
            a[i] = q[i * SW];
            m[i] = fminf(q[i * SW - 1], q[i * SW + 1]);
          } else {
            a[i] = INFINITY; m[i] = INFINITY;
          }
        }
#pragma unroll
        for (int j = 0; j < kDtcVS; ++j) {
          if (j < nrow[u]) {
            const float e = fminf(m[j + 1], fminf(a[j], a[j + 2]));
            const float x = fminf(m[j], m[j + 2]);
            float best = fminf(old[cbase[u] + j * nw], __fadd_rn(a[j + 1], l00));
            best = fminf(best, fminf(__fadd_rn(e, l01), __fadd_rn(x, l11)));
            out[sbase[u] + j * SW] = best;
            if (talk) send(nxt, u, j, best);
            dst[gbase[u] + j * sth] = best;
          }
        }
      }
    }
    if (s + kDtcPF < count) prefetch(slot, p + kDtcPF * dir);
    cp_async_commit();
    if (dbg) { const long long t = clock64(); t_comp += t - t0; t0 = t; }
    __syncthreads();                              // my rows of generation s+1 are complete; generation s is no longer read
    if (dbg) { const long long t = clock64(); t_sync += t - t0; t0 = t; }
    cur = nxt;
  }
  if (dbg && (tid & 255) == 0) {                   // four probes per CTA: cycles in each part of a step, summed over the pass
    long long* o = dbg + ((size_t)rank * 4 + (tid >> 8)) * 4;
    o[0] = t_cp; o[1] = t_halo; o[2] = t_comp; o[3] = t_sync;
  }
  umma::cluster_sync();                           // no CTA leaves while a neighbour may still address its shared memory
}

// ------------------------------------------------------------------------------ MLP pieces outside the GEMMs
// Activations h_l and back-propagated deltas are kept ROW-MAJOR only ([points][128 features], split planes).  The
// forward and dX GEMMs read them as K-major operands (K = features); the weight-gradient kernel k_nsf_dw reads the
// SAME tiles as MN-major operands (M / N = features, K = points) -- the 64-byte-swizzled tile TMA writes is both at
// once (scripts/exp/mn_major.cu, profiles/r02_exp_mn_major_operands.txt).  Round 1 kept a transposed copy of every
// tensor for the dW GEMMs: 2-byte scattered stores in 15 epilogues and twice the activation traffic.
struct NsfBufs {
  const float4* x4;                // [n_pad] source points (xyz, 0)
  __nv_bfloat16* x16;              // [planes][n_pad][32]: (x, y, z, 1, 0...) per point, 0 rows beyond n -- the "h_0" of k_nsf_dw
  int n, n_pad;
  int planes;
  long long ps;                    // plane stride of the [n_pad][128] tensors
  __nv_bfloat16* H[kNsfLayers + 1];    // H[l], l = 1..8: activations
  __nv_bfloat16* DL[kNsfLayers + 1];   // DL[l], l = 1..8: delta_l = d loss / d (pre-activation of h_l), scaled by grad_scale
  float* params;                   // master fp32 parameters, reference layout
  float* flow;                     // [n_pad][4]
  float* best_flow;                // [n_pad][4]
  float* head_part;                // [head_blocks][kHeadPart]
  float* gW0;                      // [128][3] + [128] bias
  float* gb;                       // [7][128] bias grads of layers 1..7 (scaled)
  float* small_part;               // [9][splits][128][4]: per split (delta_{l+1}^T x, delta_{l+1}^T 1) for l = 0..7, h_8^T dflow for l = 8
  __nv_bfloat16* d16;              // [planes][n_pad][32]: (dflow_x, dflow_y, dflow_z, 0...) per point, scaled like the deltas
  uint32_t* relu_mask[kNsfLayers]; // relu_mask[l], l = 1..7: [n_pad][4] bit j of word c = (h_l[32 c + j] > 0), written by the fused
                                   // forward chain and read by the fused dX chain instead of the 50 MB activation tensor
  NsfCtl* ctl;
  float grad_scale;                // power of two S; every backward tensor carries S/N instead of 1/N
};

// parameter offsets in the flat master array (reference state_dict order)
__host__ __device__ inline int nsf_off_w(int l) {   // l = 0..8
  if (l == 0) return 0;
  return 128 * 3 + 128 + (l - 1) * (128 * 128 + 128);
}
__host__ __device__ inline int nsf_off_b(int l) {
  if (l == 0) return 128 * 3;
  if (l == 8) return nsf_off_w(8) + 3 * 128;
  return nsf_off_w(l) + 128 * 128;
}
constexpr int kNsfParams = 128 * 3 + 128 + 7 * (128 * 128 + 128) + 3 * 128 + 3;   // 116483

// layer 0: h1 = relu(W0 x + b0), K = 3 -> plain FMAs.  A thread owns 32 consecutive features of one point
// (64-byte row-major stores per plane).
constexpr int kL0Pts = 64;
__global__ void __launch_bounds__(256)
k_nsf_l0_fwd(NsfBufs b) {
  if (b.ctl->stop) return;
  __shared__ float sW[384], sB[128];
  for (int t = threadIdx.x; t < 384; t += 256) sW[t] = b.params[nsf_off_w(0) + t];
  for (int t = threadIdx.x; t < 128; t += 256) sB[t] = b.params[nsf_off_b(0) + t];
  __syncthreads();
  const bool split = b.planes == 2;
  for (long long base = (long long)blockIdx.x * kL0Pts; base < b.n_pad; base += (long long)gridDim.x * kL0Pts) {
    const int pi = threadIdx.x >> 2, j0 = (threadIdx.x & 3) * 32;
    const long long i = base + pi;
    const float4 p = b.x4[i];
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = j0 + 2 * k + u;
        float t = sW[j * 3] * p.x;
        t = fmaf(sW[j * 3 + 1], p.y, t);
        t = fmaf(sW[j * 3 + 2], p.z, t);
        v[u] = fmaxf(t + sB[j], 0.f);
      }
      umma::pack_split2(v[0], v[1], split, hi[k], lo[k]);
    }
    uint4* d0 = (uint4*)(b.H[1] + i * 128 + j0);
#pragma unroll
    for (int k = 0; k < 4; ++k) d0[k] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
    if (split) {
      uint4* d1 = (uint4*)(b.H[1] + b.ps + i * 128 + j0);
#pragma unroll
      for (int k = 0; k < 4; ++k) d1[k] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
    }
  }
}

// 8 consecutive features of one point from a row-major split-plane tensor
__device__ __forceinline__ void load_split8(const __nv_bfloat16* p, long long ps, int planes, float v[8]) {
  const uint4 a = *(const uint4*)p;
  const uint32_t aa[4] = {a.x, a.y, a.z, a.w};
  if (planes == 2) {
    const uint4 l = *(const uint4*)(p + ps);
    const uint32_t ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 h2 = __half22float2(*reinterpret_cast<const __half2*>(&aa[k]));
      const float2 l2 = __half22float2(*reinterpret_cast<const __half2*>(&ll[k]));
      v[2 * k] = h2.x + l2.x; v[2 * k + 1] = h2.y + l2.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[2 * k] = __uint_as_float(aa[k] << 16); v[2 * k + 1] = __uint_as_float(aa[k] & 0xffff0000u); }
  }
}

constexpr int kHeadThreads = 128;
constexpr int kHeadPart = 4;                 // per-block partial sums: db8 [3], loss

// output layer of point i: flow = W8 h8 + b8 (W8 staged in shared memory)
__device__ __forceinline__ void head_flow(const NsfBufs& b, const float* sW8, int i, float f[3]) {
  const float* b8 = b.params + nsf_off_b(8);
  float f0 = 0.f, f1 = 0.f, f2 = 0.f;
  const __nv_bfloat16* h8 = b.H[8] + (long long)i * 128;
  for (int j0 = 0; j0 < 128; j0 += 8) {
    float h[8];
    load_split8(h8 + j0, b.ps, b.planes, h);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      f0 = fmaf(sW8[j0 + u], h[u], f0);
      f1 = fmaf(sW8[128 + j0 + u], h[u], f1);
      f2 = fmaf(sW8[256 + j0 + u], h[u], f2);
    }
  }
  f[0] = f0 + __ldg(b8); f[1] = f1 + __ldg(b8 + 1); f[2] = f2 + __ldg(b8 + 2);
}

// backward through the output layer and the ReLU of h8: delta_8 = (d W8) * relu'(h8) -> DL[8]; d -> d16 (dW8 = d^T h8 is
// then one more item of k_nsf_dw: the 384 per-block shuffle reductions this kernel used to do took 60 of its 95 us);
// per-block partial sums of db8 [3] and `extra` (the loss for FastNSF) in head_part (fixed shuffle tree => deterministic)
__device__ __forceinline__ void head_backward(const NsfBufs& b, const float* sW8, int i, float d0, float d1, float d2,
                                              float extra, float (*red)[4]) {
  float* part = b.head_part + (size_t)blockIdx.x * kHeadPart;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool in_pad = i < b.n_pad;
  const bool split = b.planes == 2;
  const __nv_bfloat16* h8 = b.H[8] + (long long)i * 128;
  if (in_pad) {
    uint32_t hi[2], lo[2];
    umma::pack_split2(d0, d1, split, hi[0], lo[0]);
    umma::pack_split2(d2, 0.f, split, hi[1], lo[1]);
    uint4* q0 = (uint4*)(b.d16 + (size_t)i * 32);
    q0[0] = make_uint4(hi[0], hi[1], 0u, 0u); q0[1] = make_uint4(0u, 0u, 0u, 0u);
    q0[2] = make_uint4(0u, 0u, 0u, 0u); q0[3] = make_uint4(0u, 0u, 0u, 0u);
    if (split) {
      uint4* q1 = (uint4*)(b.d16 + (size_t)b.n_pad * 32 + (size_t)i * 32);
      q1[0] = make_uint4(lo[0], lo[1], 0u, 0u); q1[1] = make_uint4(0u, 0u, 0u, 0u);
      q1[2] = make_uint4(0u, 0u, 0u, 0u); q1[3] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int j0 = 0; j0 < 128; j0 += 8) {
      float h[8], dhv[8];
      load_split8(h8 + j0, b.ps, b.planes, h);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        const float dh = fmaf(d2, sW8[256 + j], fmaf(d1, sW8[128 + j], d0 * sW8[j]));
        dhv[u] = h[u] > 0.f ? dh : 0.f;
      }
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) umma::pack_split2(dhv[2 * u], dhv[2 * u + 1], split, ph[u], pl[u]);
      *(uint4*)(b.DL[8] + (long long)i * 128 + j0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      if (split) *(uint4*)(b.DL[8] + b.ps + (long long)i * 128 + j0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
  float a0 = d0, a1 = d1, a2 = d2, a3 = extra;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    a2 += __shfl_xor_sync(0xffffffffu, a2, s);
    a3 += __shfl_xor_sync(0xffffffffu, a3, s);
  }
  if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = a2; red[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int wv = 0; wv < kHeadThreads / 32; ++wv) s += red[wv][threadIdx.x];
    part[threadIdx.x] = s;
  }
}

// FastNSF head: output layer + loss + its backward, one thread per point:
//   flow = W8 h8 + b8; Y = x + flow; loss_i = trilinear D(Y); dflow = grad_Y D * (S/N); delta_8, dW8, db8, loss partials.
__global__ void __launch_bounds__(kHeadThreads)
k_nsf_head(NsfBufs b, const float* __restrict__ D, NsfVol vol) {
  if (b.ctl->stop) return;
  __shared__ float red[kHeadThreads / 32][4];
  __shared__ float sW8[384];
  for (int t = threadIdx.x; t < 384; t += kHeadThreads) sW8[t] = b.params[nsf_off_w(8) + t];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < b.n;
  const int ii = live ? i : 0;
  float f[3];
  head_flow(b, sW8, ii, f);
  const float f0 = f[0], f1 = f[1], f2 = f[2];
  const float4 x = b.x4[ii];
  const float Y[3] = {x.x + f0, x.y + f1, x.z + f2};
  // DT.torch_bilinear_distance (fastnsf.py:59-80) followed by grid_sample's own un-normalisation
  float ix[3], pass[3];
  int i0[3];
  float fr[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float sz = (float)(vol.dims[k] - 1);
    const float raw = __fmul_rn(__fsub_rn(Y[k], vol.lo[k]), vol.gf);
    const float s = fminf(fmaxf(raw, 0.f), sz);
    pass[k] = (raw >= 0.f && raw <= sz) ? vol.gf : 0.f;          // d clip / dY
    const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, s), sz), 1.f);
    ix[k] = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), sz);
    const float fl = floorf(ix[k]);
    i0[k] = (int)fl;
    fr[k] = ix[k] - fl;
  }
  float val = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int dx = c >> 2, dy = (c >> 1) & 1, dz = c & 1;
    const int cx = i0[0] + dx, cy = i0[1] + dy, cz = i0[2] + dz;
    float dv = 0.f;   // zero padding outside the volume
    if (cx >= 0 && cx < vol.dims[0] && cy >= 0 && cy < vol.dims[1] && cz >= 0 && cz < vol.dims[2])
      dv = __ldg(D + ((size_t)cx * vol.dims[1] + cy) * vol.dims[2] + cz);
    const float wx = dx ? fr[0] : 1.f - fr[0], wy = dy ? fr[1] : 1.f - fr[1], wz = dz ? fr[2] : 1.f - fr[2];
    val = fmaf(dv, wx * wy * wz, val);
    gx = fmaf(dv, (dx ? 1.f : -1.f) * wy * wz, gx);
    gy = fmaf(dv, (dy ? 1.f : -1.f) * wx * wz, gy);
    gz = fmaf(dv, (dz ? 1.f : -1.f) * wx * wy, gz);
  }
  // d ix / dY = (sz/2) * (2/sz) * gf * [inside] = gf * [inside]
  const float sN = b.grad_scale / (float)b.n;
  const float d0 = live ? gx * pass[0] * sN : 0.f, d1 = live ? gy * pass[1] * sN : 0.f, d2 = live ? gz * pass[2] * sN : 0.f;
  if (!live) val = 0.f;
  if (live) *(float4*)(b.flow + 4 * (size_t)i) = make_float4(f0, f1, f2, val);
  head_backward(b, sW8, i, d0, d1, d2, val, red);
}

// The same head with a WARP per point instead of a thread per point (himo_nsf_set_head_warp, default): lanes own 4 features
// each, so h_8 is read and delta_8 written as coalesced 256-byte rows (a thread per point walked its own 512 bytes: 32 lines
// per load instruction), the three dot products are shuffle-reduced, and four points go through each stage together so that
// the dependent global loads (h_8, then the eight distance-volume corners, fetched by eight lanes per point) are in flight
// for four points at once.  Block = 128 points as before (same head_part layout); sums are taken in a fixed order.
__global__ void __launch_bounds__(kHeadThreads)
k_nsf_head_warp(NsfBufs b, const float* __restrict__ D, NsfVol vol) {
  if (b.ctl->stop) return;
  __shared__ float red[kHeadThreads / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool split = b.planes == 2;
  const unsigned FULL = 0xffffffffu;
  float w[3][4];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int u = 0; u < 4; ++u) w[k][u] = __ldg(b.params + nsf_off_w(8) + k * 128 + lane * 4 + u);
  const float b8x = __ldg(b.params + nsf_off_b(8)), b8y = __ldg(b.params + nsf_off_b(8) + 1), b8z = __ldg(b.params + nsf_off_b(8) + 2);
  const float sN = b.grad_scale / (float)b.n;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int base = blockIdx.x * kHeadThreads + warp * 32;
  const int q = lane >> 3, corner = lane & 7;           // trilinear stage: 8 lanes per point, one corner each
  for (int p0 = 0; p0 < 32; p0 += 4) {
    const int i0 = base + p0;
    if (i0 >= b.n_pad) break;                            // n_pad is a multiple of 256: groups of four are whole
    // ---- h_8 rows of four points, 4 features per lane
    float h[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const __nv_bfloat16* row = b.H[8] + (long long)(i0 + t) * 128 + lane * 4;
      const uint2 a = *(const uint2*)row;
      if (split) {
        const uint2 l = *(const uint2*)(row + b.ps);
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), h1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
        const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
        h[t][0] = h0.x + l0.x; h[t][1] = h0.y + l0.y; h[t][2] = h1.x + l1.x; h[t][3] = h1.y + l1.y;
      } else {
        h[t][0] = __uint_as_float(a.x << 16); h[t][1] = __uint_as_float(a.x & 0xffff0000u);
        h[t][2] = __uint_as_float(a.y << 16); h[t][3] = __uint_as_float(a.y & 0xffff0000u);
      }
    }
    // ---- flow = W8 h8 + b8: per-lane partial dot products, xor tree (every lane ends with the sums)
    float f[4][3];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float acc = w[k][0] * h[t][0];
        acc = fmaf(w[k][1], h[t][1], acc); acc = fmaf(w[k][2], h[t][2], acc); acc = fmaf(w[k][3], h[t][3], acc);
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(FULL, acc, sft);
        f[t][k] = acc;
      }
    // ---- trilinear lookup of point i0 + q, corner `corner`
    const int i = i0 + q;
    const bool live = i < b.n;
    const float f0 = (q == 0 ? f[0][0] : q == 1 ? f[1][0] : q == 2 ? f[2][0] : f[3][0]) + b8x;
    const float f1 = (q == 0 ? f[0][1] : q == 1 ? f[1][1] : q == 2 ? f[2][1] : f[3][1]) + b8y;
    const float f2 = (q == 0 ? f[0][2] : q == 1 ? f[1][2] : q == 2 ? f[2][2] : f[3][2]) + b8z;
    const float4 x = b.x4[i];
    const float Y[3] = {x.x + f0, x.y + f1, x.z + f2};
    // DT.torch_bilinear_distance (fastnsf.py:59-80) followed by grid_sample's own un-normalisation
    float pass[3], fr[3];
    int c0[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float sz = (float)(vol.dims[k] - 1);
      const float raw = __fmul_rn(__fsub_rn(Y[k], vol.lo[k]), vol.gf);
      const float sc = fminf(fmaxf(raw, 0.f), sz);
      pass[k] = (raw >= 0.f && raw <= sz) ? vol.gf : 0.f;          // d clip / dY
      const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, sc), sz), 1.f);
      const float ixk = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), sz);
      const float fl = floorf(ixk);
      c0[k] = (int)fl;
      fr[k] = ixk - fl;
    }
    const int dx = corner >> 2, dy = (corner >> 1) & 1, dz = corner & 1;
    const int cx = c0[0] + dx, cy = c0[1] + dy, cz = c0[2] + dz;
    float dv = 0.f;   // zero padding outside the volume
    if (cx >= 0 && cx < vol.dims[0] && cy >= 0 && cy < vol.dims[1] && cz >= 0 && cz < vol.dims[2])
      dv = __ldg(D + ((size_t)cx * vol.dims[1] + cy) * vol.dims[2] + cz);
    const float wx = dx ? fr[0] : 1.f - fr[0], wy = dy ? fr[1] : 1.f - fr[1], wz = dz ? fr[2] : 1.f - fr[2];
    float val = dv * (wx * wy * wz);
    float gx = dv * ((dx ? 1.f : -1.f) * wy * wz), gy = dv * ((dy ? 1.f : -1.f) * wx * wz), gz = dv * ((dz ? 1.f : -1.f) * wx * wy);
#pragma unroll
    for (int sft = 4; sft > 0; sft >>= 1) {               // the 8 corners of a point sit in 8 adjacent lanes
      val += __shfl_xor_sync(FULL, val, sft); gx += __shfl_xor_sync(FULL, gx, sft);
      gy += __shfl_xor_sync(FULL, gy, sft); gz += __shfl_xor_sync(FULL, gz, sft);
    }
    // d ix / dY = (sz/2) * (2/sz) * gf * [inside] = gf * [inside]
    const float d0 = live ? gx * pass[0] * sN : 0.f, d1 = live ? gy * pass[1] * sN : 0.f, d2 = live ? gz * pass[2] * sN : 0.f;
    if (!live) val = 0.f;
    if (live && corner == 0) *(float4*)(b.flow + 4 * (size_t)i) = make_float4(f0, f1, f2, val);
    // ---- back through the output layer: every lane needs (d, val) of all four points
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float e0 = __shfl_sync(FULL, d0, t * 8), e1 = __shfl_sync(FULL, d1, t * 8), e2 = __shfl_sync(FULL, d2, t * 8);
      const float ev = __shfl_sync(FULL, val, t * 8);
      a0 += e0; a1 += e1; a2 += e2; a3 += ev;
      const long long it = i0 + t;
      float dh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float g = fmaf(e2, w[2][u], fmaf(e1, w[1][u], e0 * w[0][u]));
        dh[u] = h[t][u] > 0.f ? g : 0.f;
      }
      uint32_t ph[2], pl[2];
      umma::pack_split2(dh[0], dh[1], split, ph[0], pl[0]);
      umma::pack_split2(dh[2], dh[3], split, ph[1], pl[1]);
      *(uint2*)(b.DL[8] + it * 128 + lane * 4) = make_uint2(ph[0], ph[1]);
      if (split) *(uint2*)(b.DL[8] + b.ps + it * 128 + lane * 4) = make_uint2(pl[0], pl[1]);
      // d16 row (dflow_x, dflow_y, dflow_z, 0...): 64 bytes per plane, lanes 0..3 the hi row, 4..7 the lo row
      if (lane < (split ? 8 : 4)) {
        uint32_t hi[2], lo[2];
        umma::pack_split2(e0, e1, split, hi[0], lo[0]);
        umma::pack_split2(e2, 0.f, split, hi[1], lo[1]);
        const bool lo_row = lane >= 4;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if ((lane & 3) == 0) v = lo_row ? make_uint4(lo[0], lo[1], 0u, 0u) : make_uint4(hi[0], hi[1], 0u, 0u);
        *((uint4*)(b.d16 + (lo_row ? (size_t)b.n_pad * 32 : 0) + (size_t)it * 32) + (lane & 3)) = v;
      }
    }
  }
  // per-block partial sums of db8 [3] and the loss (a* are identical in every lane of a warp)
  if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = a2; red[warp][3] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float sm = 0.f;
    for (int wv = 0; wv < kHeadThreads / 32; ++wv) sm += red[wv][threadIdx.x];
    b.head_part[(size_t)blockIdx.x * kHeadPart + threadIdx.x] = sm;
  }
}

// loss reduction + best-flow bookkeeping + EarlyStopping.step (nsfp_module.py:65-82), single thread.
__global__ void k_nsf_control(NsfBufs b, int head_blocks, float min_delta, int patience) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  NsfCtl* c = b.ctl;
  const int stopped = c->stop;
  double s = 0.0;
  if (!stopped)   // 32 strided partial sums + a fixed shuffle tree (deterministic)
    for (int k = threadIdx.x; k < head_blocks; k += 32) s += (double)b.head_part[(size_t)k * kHeadPart + 3];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (threadIdx.x != 0) return;
  if (stopped) { c->snapshot = 0; return; }
  const float loss = (float)(s / (double)b.n);
  c->loss = loss;
  c->iters += 1;
  c->snapshot = 0;
  if (loss <= c->best_loss) { c->best_loss = loss; c->snapshot = 1; }     // fastnsf.py:151-153
  int stop = 0;
  if (patience == 0) stop = 0;   // EarlyStopping(patience=0): `self.step = lambda a: False` (nsfp_module.py:60-62), not even NaN stops it
  else if (!c->es_has_best) { c->es_has_best = 1; c->es_best = loss; }
  else if (isnan(loss)) stop = 1;
  else {
    if (loss < c->es_best - min_delta) { c->es_bad = 0; c->es_best = loss; }
    else c->es_bad += 1;
    if (c->es_bad >= patience) stop = 1;
  }
  c->stop = stop;
}

__global__ void __launch_bounds__(256)
k_nsf_snapshot(NsfBufs b) {
  if (!b.ctl->snapshot) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += gridDim.x * blockDim.x)
    *(float4*)(b.best_flow + 4 * (size_t)i) = *(const float4*)(b.flow + 4 * (size_t)i);
}

// ------------------------------------------------------------------------------ weight gradients: k_nsf_dw
// ALL parameter gradients of the hidden layers in ONE launch:
//   dW_l = delta_{l+1}^T h_l (l = 1..7),  db_l = delta_{l+1}^T 1 (l = 0..7),  dW_0 = delta_1^T x
// as split-K GEMMs over the points.  CTA c owns the point range [c * k_split, (c+1) * k_split) and walks the eight
// layers; per layer it accumulates in tensor memory (M = 128 out features of delta_{l+1}, N = 128 in features of h_l,
// plus a 16-column accumulator against the (x, y, z, 1) tile) and flushes its partial to dW_part / small_part, which
// k_nsf_small_final / k_nsf_adam add up over the CTAs.  Operands are the ROW-MAJOR tiles of delta and h read as
// MN-major (see NsfBufs): stage = 32 points: delta tile 4 chunks x 2 planes x 2 KB, h tile the same, x tile 2 x 2 KB.
// Split fp16 planes -> three MMAs per product (hi*hi, hi*lo, lo*hi) into ONE accumulator (<= 130 MMAs per chain).
// Accumulators are double-buffered by layer parity so that the flush of layer l overlaps the MMAs of layer l+1.
// Memory-bound by construction (each stage's 36 KB feed 12 small MMAs): 8 x 100 MB of operands per iteration.
constexpr int kDwKT = 32;                              // points per stage
constexpr int kDwChunkB = kDwKT * 64;                  // one 32-feature chunk of one plane: 2 KB
constexpr int kDwOpB = 4 * 2 * kDwChunkB;              // delta or h tile: 16 KB
constexpr int kDwStageB = 2 * kDwOpB + 2 * kDwChunkB;  // 36 KB
constexpr int kDwStages = 5;
constexpr int kDwBarOffset = kDwStages * kDwStageB;
constexpr int kDwTotal = kDwBarOffset + 256 + 1024;
constexpr int kDwThreads = 192;                        // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct DwMaps { CUtensorMap delta[kNsfLayers + 1]; CUtensorMap h[kNsfLayers + 1]; CUtensorMap x; CUtensorMap d; };   // delta[1..8], h[1..8]

__device__ __forceinline__ uint64_t dw_desc_mn(uint32_t addr) {   // MN-major, SWIZZLE_64B: LBO = chunk stride, SBO = 8 rows
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)(kDwChunkB >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;
  return d;
}

__global__ void __launch_bounds__(kDwThreads, 1)
k_nsf_dw(const __grid_constant__ DwMaps maps, const int* __restrict__ stop_flag, int n_pad, int k_split, int splits,
         int planes, float* __restrict__ dW_part, float* __restrict__ small_part) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + kDwBarOffset);
  uint64_t* empty_bar = full_bar + kDwStages;
  uint64_t* acc_full = empty_bar + kDwStages;       // [2]
  uint64_t* acc_empty = acc_full + 2;               // [2]
  uint32_t* tmem_ptr = (uint32_t*)(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  if (*stop_flag) return;
  const int cta = blockIdx.x;
  if (cta >= splits) return;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kDwStages; ++s) { umma::mbar_init(&full_bar[s], 1); umma::mbar_init(&empty_bar[s], 1); }
    for (int k = 0; k < 2; ++k) { umma::mbar_init(&acc_full[k], 1); umma::mbar_init(&acc_empty[k], 4); }
    umma::fence_barrier_init();
  } else if (warp == 1) {
    umma::tmem_alloc(tmem_ptr, 512);
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int p0 = cta * k_split;
  const int n_stage = (min(k_split, n_pad - p0) + kDwKT - 1) / kDwKT;
  // TMEM columns: buffer k: dW accumulator [k*144, +128), small accumulator [k*144 + 128, +16)
  if (warp == 0) {
    uint32_t git = 0;
    for (int l = 0; l <= kNsfLayers; ++l) {
      for (int st = 0; st < n_stage; ++st, ++git) {
        const int s = git % kDwStages;
        umma::mbar_wait(&empty_bar[s], ((git / kDwStages) & 1) ^ 1);
        if (umma::elect_one()) {
          uint8_t* dst = smem + s * kDwStageB;
          const int pt = p0 + st * kDwKT;
          // item l < 8: A = delta_{l+1}, B = h_l (l > 0), small B = (x, y, z, 1); item 8: A = h_8, small B = dflow (dW_8^T)
          const bool big = l > 0 && l < kNsfLayers;
          const CUtensorMap* mA = l < kNsfLayers ? &maps.delta[l + 1] : &maps.h[kNsfLayers];
          const CUtensorMap* mS = l < kNsfLayers ? &maps.x : &maps.d;
          umma::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(planes * (4 * kDwChunkB * (big ? 2 : 1) + kDwChunkB)));
          for (int pl = 0; pl < planes; ++pl)
#pragma unroll
            for (int c = 0; c < 4; ++c)
              umma::tma_load_3d(dst + (pl * 4 + c) * kDwChunkB, mA, &full_bar[s], c * 32, pt, pl);
          if (big) {
            for (int pl = 0; pl < planes; ++pl)
#pragma unroll
              for (int c = 0; c < 4; ++c)
                umma::tma_load_3d(dst + kDwOpB + (pl * 4 + c) * kDwChunkB, &maps.h[l], &full_bar[s], c * 32, pt, pl);
          }
          for (int pl = 0; pl < planes; ++pl)
            umma::tma_load_3d(dst + 2 * kDwOpB + pl * kDwChunkB, mS, &full_bar[s], 0, pt, pl);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t kMN = (1u << 15) | (1u << 16);
    const uint32_t fmt = planes == 2 ? 0u : 1u;          // split fp16 planes, or one bf16 plane
    const uint32_t id_big = umma::idesc_f16kind_f32(128, 128, fmt, fmt) | kMN;
    const uint32_t id_small = umma::idesc_f16kind_f32(128, 16, fmt, fmt) | kMN;
    const bool split = planes == 2;
    uint32_t git = 0;
    for (int l = 0; l <= kNsfLayers; ++l) {
      const int buf = l & 1;
      umma::mbar_wait(&acc_empty[buf], ((l >> 1) & 1) ^ 1);
      umma::tc_fence_after();
      const uint32_t t_big = tmem_base + (uint32_t)(buf * 144), t_small = t_big + 128u;
      for (int st = 0; st < n_stage; ++st, ++git) {
        const int s = git % kDwStages;
        umma::mbar_wait(&full_bar[s], (git / kDwStages) & 1);
        umma::tc_fence_after();
        if (umma::elect_one()) {
          const uint32_t a0 = umma::smem_u32(smem + s * kDwStageB), b0 = a0 + kDwOpB, x0 = a0 + 2 * kDwOpB;
#pragma unroll
          for (int k = 0; k < kDwKT / 16; ++k) {
            const uint32_t ko = (uint32_t)(k * 16 * 64);
            const uint64_t a_hi = dw_desc_mn(a0 + ko), a_lo = dw_desc_mn(a0 + 4 * kDwChunkB + ko);
            const uint32_t first = (st == 0 && k == 0) ? 0u : 1u;
            if (l > 0 && l < kNsfLayers) {
              const uint64_t b_hi = dw_desc_mn(b0 + ko), b_lo = dw_desc_mn(b0 + 4 * kDwChunkB + ko);
              umma::mma_bf16_ss(t_big, a_hi, b_hi, id_big, first);
              if (split) {
                umma::mma_bf16_ss(t_big, a_hi, b_lo, id_big, 1u);
                umma::mma_bf16_ss(t_big, a_lo, b_hi, id_big, 1u);
              }
            }
            const uint64_t x_hi = dw_desc_mn(x0 + ko), x_lo = dw_desc_mn(x0 + kDwChunkB + ko);
            umma::mma_bf16_ss(t_small, a_hi, x_hi, id_small, first);
            if (split) {
              umma::mma_bf16_ss(t_small, a_hi, x_lo, id_small, 1u);
              umma::mma_bf16_ss(t_small, a_lo, x_hi, id_small, 1u);
            }
          }
          umma::mma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (umma::elect_one()) umma::mma_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;                 // out feature (row of dW)
    for (int l = 0; l <= kNsfLayers; ++l) {
      const int buf = l & 1;
      umma::mbar_wait(&acc_full[buf], (l >> 1) & 1);
      umma::tc_fence_after();
      const uint32_t t_big = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 144);
      if (l > 0 && l < kNsfLayers) {
        float* dst = dW_part + ((size_t)(l - 1) * splits + cta) * 16384 + (size_t)m * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          umma::tmem_ld_32x32(t_big + (uint32_t)(c * 32), r);
          umma::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ((float4*)(dst + c * 32))[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        }
      }
      {
        uint32_t r[16];
        umma::tmem_ld_32x16(t_big + 128u, r);
        umma::tmem_ld_wait();
        *(float4*)(small_part + (((size_t)l * splits + cta) * 128 + m) * 4) =
            make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&acc_empty[buf]);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------ fused layer chains: k_mlp_chain
// The seven 128x128 hidden layers as ONE persistent kernel per direction instead of seven GEMM launches.  Unfused, every
// layer moved its 128-point tile through the SM three times (64 KB of activations + 64 KB of weights in, 64 KB out, for
// 768 cycles of MMA time: bound by SM<->L2 bytes).  Here a CTA pair (cta_group::2) owns 256 points for the
// whole chain, exactly like the fused ConvGRU decoder (csrc/decfused.cu):
//   * the activation tile lives in shared memory as split-fp16 planes in the K-major 64-byte-swizzled layout tcgen05.mma
//     reads (4 chunks x 2 planes x 8 KB); the epilogue warps rewrite it in place with the next layer's activations;
//   * the same tile is handed to TMA as-is (cp.async.bulk.tensor store, 8 per layer) to put h_{l+1} / delta_l into HBM for
//     the weight-gradient kernel -- no per-thread global stores;
//   * weights stream through an 8-stage TMA ring (one 32-channel k-chunk per stage, each CTA half of the 128 rows);
//   * DIR 0 (forward): prologue = layer 0 on CUDA cores (h_1 = relu(W0 x + b0)); epilogue = bias + ReLU, and one bit per
//     activation goes to relu_mask[l] (16 bytes per point per layer);
//   * DIR 1 (backward, dX): prologue = TMA load of delta_8; layers 7..1 with the transposed weights; epilogue = the ReLU
//     mask read back from those bits (instead of re-reading the 50 MB activation tensor per layer).
// One accumulator (hi*hi and cross terms together: chains are 24 MMAs), 128 TMEM columns.
constexpr int kChThreads = 352;                       // TMA producer, MMA issuer, 8 epilogue warps, store warp
constexpr int kChTile = 128 * 64;                      // one 32-channel plane tile of 128 rows: 8 KB
constexpr int kChT = 4 * 2 * kChTile;                  // activation tile: 64 KB
constexpr int kChSlots = 2;                            // two point tiles in flight per CTA: the epilogue of one overlaps the MMAs of the other
constexpr int kChWStage = 2 * 64 * 64;                 // one k-chunk of one layer, this CTA's 64 rows, both planes: 8 KB
constexpr int kChWStages = 8;
constexpr int kChOffW = kChSlots * kChT;
constexpr int kChOffBar = kChOffW + kChWStages * kChWStage;
constexpr int kChOffConst = kChOffBar + 256;
constexpr int kChConstFloats = 7 * 128 + 384 + 128;
constexpr int kChTotal = kChOffConst + kChConstFloats * 4 + 1024;

struct ChainMaps { CUtensorMap w[kNsfLayers]; CUtensorMap act[kNsfLayers + 1]; };   // w[1..7]; act[l] = H[l] (DIR 0) / DL[l] (DIR 1)

struct ChainParams {
  int n_pair_tiles;
  const float* params;             // fp32 master parameters (biases, W0)
  const float4* x4;
  uint32_t* mask[kNsfLayers];      // relu_mask[1..7]
  float acc_scale;
  const int* stop_flag;
};

__device__ __forceinline__ uint32_t ch_swz(int row, int j) { return (uint32_t)(row * 64 + ((j ^ ((row >> 1) & 3)) << 4)); }
__device__ __forceinline__ void ch_store16(uint8_t* tile_hi, uint8_t* tile_lo, int row, int j0, const float* v) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::pack_split2(v[u * 8 + 2 * k], v[u * 8 + 2 * k + 1], true, hi[k], lo[k]);
    *(uint4*)(tile_hi + ch_swz(row, j0 + u)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *(uint4*)(tile_lo + ch_swz(row, j0 + u)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Work unit = a "super-tile" of two consecutive 256-point pair tiles (slot 0, slot 1).  Layer by layer the MMA warp runs
// slot 0 then slot 1 against the SAME weight stages, while the epilogue warps turn slot 0's accumulator into the next
// layer's operand: the tensor pipe and the epilogue warps work on different slots at any time (with one tile the chain was
// a strict MMA -> epilogue -> MMA dependency loop: 5 us per layer step, 210 us per chain).
template <int DIR>
__global__ void __launch_bounds__(kChThreads, 1)
k_mlp_chain(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem + kChOffW;
  uint64_t* w_full = (uint64_t*)(smem + kChOffBar);      // [8]
  uint64_t* w_empty = w_full + kChWStages;                // [8]
  uint64_t* t_ready = w_empty + kChWStages;               // [2] slot's tile holds the next layer's input (2 arrivals: one per CTA)
  uint64_t* t_tma = t_ready + 2;                          // [2] DIR 1: delta_8 landed (tx, both CTAs)
  uint64_t* t_free = t_tma + 2;                           // [2] DIR 1: this CTA's tile may be overwritten by the next TMA load
  uint64_t* acc_full = t_free + 2;                        // [2]
  uint64_t* st_req = acc_full + 2;                        // [2] slot's tile is complete: the store warp may send it to HBM
  uint64_t* st_done = st_req + 2;                         // [2] that store has read the tile: it may be rewritten
  uint32_t* tmem_ptr_smem = (uint32_t*)(st_done + 2);
  float* c_bias = (float*)(smem + kChOffConst);          // [7][128] biases of layers 1..7
  float* c_w0 = c_bias + 7 * 128;                         // [128][3]
  float* c_b0 = c_w0 + 384;                               // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (*p.stop_flag) return;
  const uint32_t cta_rank = umma::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_workers = (int)(gridDim.x >> 1), worker = (int)(blockIdx.x >> 1);
  const int n_super = (p.n_pair_tiles + 1) >> 1;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kChWStages; ++s) { umma::mbar_init(&w_full[s], 1); umma::mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(&t_ready[s], 2); umma::mbar_init(&t_tma[s], 1); umma::mbar_init(&t_free[s], 1); umma::mbar_init(&acc_full[s], 1);
      umma::mbar_init(&st_req[s], 1); umma::mbar_init(&st_done[s], 1);
    }
    umma::fence_barrier_init();
  }
  if (DIR == 0) {
    for (int i = threadIdx.x; i < 7 * 128; i += kChThreads) {
      const int l = 1 + i / 128;
      c_bias[i] = p.params[nsf_off_b(l) + (i & 127)];
    }
    for (int i = threadIdx.x; i < 384; i += kChThreads) c_w0[i] = p.params[nsf_off_w(0) + i];
    for (int i = threadIdx.x; i < 128; i += kChThreads) c_b0[i] = p.params[nsf_off_b(0) + i];
  }
  __syncthreads();
  umma::cluster_sync();
  // tcgen05.alloc.cta_group::2 is a compiler-generated handshake through the PEER CTA's reserved shared memory (remote
  // mbarrier arrive + remote store of the address): it may only run once the peer CTA is known to be executing, i.e. after a
  // cluster barrier.  Allocating before it hangs when the two CTAs of a pair start far apart, which several streams' kernels
  // sharing the GPU provoke (profiles/r02_two_cta_alloc_hang.txt).
  if (warp == 1) umma::tmem_alloc_2cta(tmem_ptr_smem, 256);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // layer order: DIR 0: l = 1..7 (input h_l, output h_{l+1});  DIR 1: l = 7..1 (input delta_{l+1}, output delta_l)
  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t git = 0;
    int uc = 0;
    for (int u = worker; u < n_super; u += n_workers, ++uc) {
      const int nslot = (2 * u + 1 < p.n_pair_tiles) ? 2 : 1;
      if (DIR == 1) {
        for (int sl = 0; sl < nslot; ++sl) {
          const int row0 = ((2 * u + sl) * 2 + (int)cta_rank) * 128;
          if (uc > 0) umma::mbar_wait(&t_free[sl], (uint32_t)((uc - 1) & 1));     // own tile: last stores have read it
          const uint32_t fb = umma::mapa_u32(umma::smem_u32(&t_tma[sl]), 0);
          uint8_t* sT = smem + sl * kChT;
          if (umma::elect_one()) {
            if (leader) umma::mbar_arrive_expect_tx(&t_tma[sl], (uint32_t)(2 * kChT));
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int pl = 0; pl < 2; ++pl)
                umma::tma_load_3d_2cta(sT + (c * 2 + pl) * kChTile, &maps.act[kNsfLayers], fb, c * 32, row0, pl);
          }
          __syncwarp();
        }
      }
      for (int k = 0; k < 7; ++k) {
        const int l = DIR == 0 ? 1 + k : 7 - k;
        for (int c = 0; c < 4; ++c, ++git) {
          const int s = git % kChWStages;
          umma::mbar_wait(&w_empty[s], ((git / kChWStages) & 1) ^ 1);
          const uint32_t fb = umma::mapa_u32(umma::smem_u32(&w_full[s]), 0);
          uint8_t* dst = sW + s * kChWStage;
          if (umma::elect_one()) {
            if (leader) umma::mbar_arrive_expect_tx(&w_full[s], (uint32_t)(2 * kChWStage));
            umma::tma_load_3d_2cta(dst, &maps.w[l], fb, c * 32, (int)cta_rank * 64, 0);
            umma::tma_load_3d_2cta(dst + 64 * 64, &maps.w[l], fb, c * 32, (int)cta_rank * 64, 1);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = umma::idesc_f16kind_f32(256, 128, 0u, 0u);
      uint32_t git = 0, nready[2] = {0u, 0u};
      int uc = 0;
      for (int u = worker; u < n_super; u += n_workers, ++uc) {
        const int nslot = (2 * u + 1 < p.n_pair_tiles) ? 2 : 1;
        for (int k = 0; k < 7; ++k, git += 4) {
          for (int sl = 0; sl < nslot; ++sl) {
            if (DIR == 1 && k == 0) umma::mbar_wait(&t_tma[sl], (uint32_t)(uc & 1));
            else { umma::mbar_wait(&t_ready[sl], nready[sl] & 1); ++nready[sl]; }
            umma::tc_fence_after();
            const uint32_t t_addr = umma::smem_u32(smem + sl * kChT);
            const uint32_t t_acc = tmem_base + (uint32_t)(sl * 128);
            for (int c = 0; c < 4; ++c) {
              const uint32_t gi = git + (uint32_t)c;
              const int s = gi % kChWStages;
              umma::mbar_wait(&w_full[s], (gi / kChWStages) & 1);
              umma::tc_fence_after();
              if (umma::elect_one()) {
                const uint32_t a_hi = t_addr + (c * 2) * kChTile, a_lo = a_hi + kChTile;
                const uint32_t b_hi = umma::smem_u32(sW + s * kChWStage), b_lo = b_hi + 64 * 64;
                const uint64_t dah = umma::smem_desc_kmajor<64>(a_hi), dal = umma::smem_desc_kmajor<64>(a_lo);
                const uint64_t dbh = umma::smem_desc_kmajor<64>(b_hi), dbl = umma::smem_desc_kmajor<64>(b_lo);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  const uint64_t koff = (uint64_t)(kk * 32 >> 4);
                  umma::mma_bf16_ss_2cta(t_acc, dah + koff, dbl + koff, idesc, (c == 0 && kk == 0) ? 0u : 1u);
                  umma::mma_bf16_ss_2cta(t_acc, dal + koff, dbh + koff, idesc, 1u);
                  umma::mma_bf16_ss_2cta(t_acc, dah + koff, dbh + koff, idesc, 1u);
                }
                if (sl == nslot - 1) umma::mma_commit_2cta(&w_empty[s]);       // the last slot has used this weight stage
              }
              __syncwarp();
            }
            if (umma::elect_one()) umma::mma_commit_2cta(&acc_full[sl]);
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 10) {
    // ===================== store warp: the HBM copy of every finished tile =====================
    // cp.async.bulk.tensor stores stall their thread while the SM's TMA queue is backed up behind DRAM.  Issued by an
    // epilogue thread they held up the epilogue (and, before the ready signal was moved in front of them, the MMAs): every
    // stored byte cost its DRAM write time on top of the compute.  Here they stall a warp that has nothing else to do.
    uint32_t nst[2] = {0u, 0u};
    auto store_tile = [&](const CUtensorMap* tm, int sl, int row0) {
      umma::mbar_wait(&st_req[sl], nst[sl] & 1);
      ++nst[sl];
      uint8_t* sT = smem + sl * kChT;
      if (umma::elect_one()) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) umma::tma_store_3d(tm, sT + (c * 2 + pl) * kChTile, c * 32, row0, pl);
        umma::tma_store_commit();
        umma::tma_store_wait_read();                       // (the elected lane owns the bulk groups)
        umma::mbar_arrive(&st_done[sl]);
      }
      __syncwarp();
    };
    for (int u = worker; u < n_super; u += n_workers) {
      const int nslot = (2 * u + 1 < p.n_pair_tiles) ? 2 : 1;
      if (DIR == 0)
        for (int sl = 0; sl < nslot; ++sl) store_tile(&maps.act[1], sl, ((2 * u + sl) * 2 + (int)cta_rank) * 128);
      for (int k = 0; k < 7; ++k) {
        const int l = DIR == 0 ? 1 + k : 7 - k;
        for (int sl = 0; sl < nslot; ++sl) {
          store_tile(&maps.act[DIR == 0 ? l + 1 : l], sl, ((2 * u + sl) * 2 + (int)cta_rank) * 128);
          // DIR 1: the tiles may be reloaded with the next super-tile's delta_8 once their last stores have read them
          if (DIR == 1 && k == 6 && sl == nslot - 1 && umma::elect_one())
            for (int z = 0; z < nslot; ++z) umma::mbar_arrive(&t_free[z]);
        }
      }
    }
    if (umma::elect_one()) umma::tma_store_wait_all();       // same lane as every elect_one above: the warp is fully converged
  } else {
    // ===================== epilogue warps: layer-0 prologue (DIR 0), then one in-place epilogue per (layer, slot) =====================
    const int q = warp & 3, half = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_t_ready0 = umma::mapa_u32(umma::smem_u32(&t_ready[0]), 0);
    const uint32_t a_t_ready1 = umma::mapa_u32(umma::smem_u32(&t_ready[1]), 0);
    const bool issuer = warp == 2 && lane == 0;
    uint32_t nacc[2] = {0u, 0u}, nreq[2] = {0u, 0u};
    // tile complete in shared memory: release it to the MMA warp (ready != 0: the leader's t_ready barrier) and to the
    // store warp
    auto publish = [&](int sl, uint32_t ready) {
      umma::tc_fence_before();
      umma::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (issuer) {
        if (ready) umma::mbar_arrive_cluster(ready);
        umma::mbar_arrive(&st_req[sl]);
      }
      ++nreq[sl];
    };
    // before a slot's tile is rewritten: its previous store (if any) has read it
    auto tile_writable = [&](int sl) { if (nreq[sl] > 0) umma::mbar_wait(&st_done[sl], (nreq[sl] - 1) & 1); };
    for (int u = worker; u < n_super; u += n_workers) {
      const int nslot = (2 * u + 1 < p.n_pair_tiles) ? 2 : 1;
      if (DIR == 0) {
        // ---- layer 0: h_1 = relu(W0 x + b0), 64 channels per thread, straight into the operand tiles
        for (int sl = 0; sl < nslot; ++sl) {
          tile_writable(sl);                                      // the previous super-tile's last store has read the tile
          uint8_t* sT = smem + sl * kChT;
          const int row0 = ((2 * u + sl) * 2 + (int)cta_rank) * 128;
          const long long grow = (long long)row0 + row;
          const float4 x = p.x4[grow];
          uint32_t mbits[2] = {0u, 0u};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = half * 64 + g * 16;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ch = col + j;
              float t = c_w0[ch * 3] * x.x;
              t = fmaf(c_w0[ch * 3 + 1], x.y, t);
              t = fmaf(c_w0[ch * 3 + 2], x.z, t);
              v[j] = fmaxf(t + c_b0[ch], 0.f);
              if (v[j] > 0.f) mbits[g >> 1] |= 1u << ((g & 1) * 16 + j);
            }
            const int c = col >> 5;
            ch_store16(sT + (c * 2) * kChTile, sT + (c * 2 + 1) * kChTile, row, (g & 1) * 2, v);
          }
          *(uint2*)(p.mask[1] + grow * 4 + half * 2) = make_uint2(mbits[0], mbits[1]);
          publish(sl, sl ? a_t_ready1 : a_t_ready0);
        }
      }
      for (int k = 0; k < 7; ++k) {
        const int l = DIR == 0 ? 1 + k : 7 - k;
        for (int sl = 0; sl < nslot; ++sl) {
          uint8_t* sT = smem + sl * kChT;
          const int row0 = ((2 * u + sl) * 2 + (int)cta_rank) * 128;
          const long long grow = (long long)row0 + row;
          umma::mbar_wait(&acc_full[sl], nacc[sl] & 1);
          ++nacc[sl];
          umma::tc_fence_after();
          uint32_t mbits[2] = {0u, 0u};
          if (DIR == 1) {
            const uint2 mm = *(const uint2*)(p.mask[l] + grow * 4 + half * 2);
            mbits[0] = mm.x; mbits[1] = mm.y;
          }
          uint32_t accs[4][16];                                   // all four loads in flight before ONE wait
#pragma unroll
          for (int g = 0; g < 4; ++g) umma::tmem_ld_32x16(tlane + (uint32_t)(sl * 128 + half * 64 + g * 16), accs[g]);
          // the store of THIS slot's previous layer has read the tile (waited for as late as possible: the accumulator
          // loads above are already on their way)
          tile_writable(sl);
          umma::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = half * 64 + g * 16;
            const uint32_t* acc = accs[g];
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (DIR == 0) {
                v[j] = fmaxf(__fmaf_rn(__uint_as_float(acc[j]), p.acc_scale, c_bias[(l - 1) * 128 + col + j]), 0.f);
                if (v[j] > 0.f) mbits[g >> 1] |= 1u << ((g & 1) * 16 + j);
              } else {
                const float t = __uint_as_float(acc[j]) * p.acc_scale;
                v[j] = ((mbits[g >> 1] >> ((g & 1) * 16 + j)) & 1u) ? t : 0.f;
              }
            }
            const int c = col >> 5;
            ch_store16(sT + (c * 2) * kChTile, sT + (c * 2 + 1) * kChTile, row, (g & 1) * 2, v);
          }
          if (DIR == 0 && l < 7) *(uint2*)(p.mask[l + 1] + grow * 4 + half * 2) = make_uint2(mbits[0], mbits[1]);
          publish(sl, k < 6 ? (sl ? a_t_ready1 : a_t_ready0) : 0u);
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync();
  if (warp == 1) umma::tmem_dealloc_2cta(tmem_base, 256);
}

// ------------------------------------------------------------------------------ generic MLP entry points (NSFP)
// The same 3 -> 128 x 8 -> 3 prior with the loss OUTSIDE: himo_mlp_forward returns the network output, the caller
// (NSFP: truncated Chamfer of two networks, OSF/src/models/nsfp.py:48-72) computes any loss and hands d loss / d output
// to himo_mlp_backward, which produces the parameter gradients (kept in the workspace for himo_mlp_adam_step) and,
// if asked, d loss / d input (NSFP differentiates through the input of its inverse network).
__global__ void __launch_bounds__(kHeadThreads)
k_mlp_head_fwd(NsfBufs b, float* __restrict__ out) {
  if (b.ctl->stop) return;
  __shared__ float sW8[384];
  for (int t = threadIdx.x; t < 384; t += kHeadThreads) sW8[t] = b.params[nsf_off_w(8) + t];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  float f[3];
  head_flow(b, sW8, i, f);
  *(float4*)(b.flow + 4 * (size_t)i) = make_float4(f[0], f[1], f[2], 0.f);
  out[3 * (size_t)i] = f[0]; out[3 * (size_t)i + 1] = f[1]; out[3 * (size_t)i + 2] = f[2];
}

// backward entry: d_out [n,3] (true scale) -> delta_8 and the dW8 / db8 partials (same reductions as k_nsf_head)
__global__ void __launch_bounds__(kHeadThreads)
k_mlp_head_bwd(NsfBufs b, const float* __restrict__ d_out) {
  if (b.ctl->stop) return;
  __shared__ float red[kHeadThreads / 32][4];
  __shared__ float sW8[384];
  for (int t = threadIdx.x; t < 384; t += kHeadThreads) sW8[t] = b.params[nsf_off_w(8) + t];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < b.n;
  const float S = b.grad_scale;
  const float d0 = live ? d_out[3 * (size_t)i] * S : 0.f, d1 = live ? d_out[3 * (size_t)i + 1] * S : 0.f,
              d2 = live ? d_out[3 * (size_t)i + 2] * S : 0.f;
  head_backward(b, sW8, i, d0, d1, d2, 0.f, red);
}

// d loss / d input = delta_1 W0 (the only path from the input: h1 = relu(W0 x + b0))
__global__ void __launch_bounds__(128)
k_mlp_dx(NsfBufs b, float inv_scale, float* __restrict__ dx) {
  if (b.ctl->stop) return;
  __shared__ float sW[384];
  for (int t = threadIdx.x; t < 384; t += 128) sW[t] = b.params[nsf_off_w(0) + t];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b.n) return;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  for (int j0 = 0; j0 < 128; j0 += 8) {
    float v[8];
    load_split8(b.DL[1] + (long long)i * 128 + j0, b.ps, b.planes, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = j0 + k;
      g0 = fmaf(v[k], sW[j * 3], g0); g1 = fmaf(v[k], sW[j * 3 + 1], g1); g2 = fmaf(v[k], sW[j * 3 + 2], g2);
    }
  }
  dx[3 * (size_t)i] = g0 * inv_scale; dx[3 * (size_t)i + 1] = g1 * inv_scale; dx[3 * (size_t)i + 2] = g2 * inv_scale;
}

// best-so-far bookkeeping + EarlyStopping.step on a loss the CALLER computed (device scalar), same state machine as
// k_nsf_control (nsfp.py:104-113, nsfp_module.py:60-82).  Counts the iteration: Adam's bias correction reads ctl->iters.
__global__ void k_mlp_control(NsfCtl* c, const float* __restrict__ loss_dev, float min_delta, int patience) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (c->stop) { c->snapshot = 0; return; }
  const float loss = *loss_dev;
  c->loss = loss;
  c->iters += 1;
  c->snapshot = 0;
  if (loss <= c->best_loss) { c->best_loss = loss; c->snapshot = 1; }
  int stop = 0;
  if (patience == 0) stop = 0;
  else if (!c->es_has_best) { c->es_has_best = 1; c->es_best = loss; }
  else if (isnan(loss)) stop = 1;
  else {
    if (loss < c->es_best - min_delta) { c->es_bad = 0; c->es_best = loss; }
    else c->es_bad += 1;
    if (c->es_bad >= patience) stop = 1;
  }
  c->stop = stop;
}
__global__ void __launch_bounds__(256)
k_mlp_snapshot(const NsfCtl* __restrict__ c, const float* __restrict__ src, float* __restrict__ dst, long long count) {
  if (!c->snapshot) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

struct NsfAdamArgs {
  float* params; float* m; float* v;
  const float* dW_part;      // [7][splits][128][128]
  int splits;
  const float* small_part;   // [9][splits][128][4] (k_nsf_dw): columns 0..2 = delta_1^T x (l = 0) / h_8^T dflow (l = 8), column 3 = bias gradient
  const float* gb;           // unused (kept for layout stability)
  const float* gW0;
  const float* gb0;
  const float* head_part; int head_blocks;
  __nv_bfloat16* Wp[kNsfLayers];    // packed [P][128][128] weights of layers 1..7 (index l)
  __nv_bfloat16* WpT[kNsfLayers];   // packed transposes
  int planes;
  float lr, beta1, beta2, eps, inv_scale;
  const NsfCtl* ctl;
};

// torch.optim.Adam.step (amsgrad=False, weight_decay=0, maximize=False) for every parameter, then the
// repack of the hidden-layer weights into the GEMM operand planes.
__global__ void __launch_bounds__(256)
k_nsf_adam(NsfAdamArgs a) {
  if (a.ctl->stop) return;
  const int step = a.ctl->iters;           // 1-based: the control kernel has already counted this iteration
  const float bc1 = 1.f - powf(a.beta1, (float)step), bc2 = 1.f - powf(a.beta2, (float)step);
  // Four lanes per parameter: lane `sub` adds the partials of the splits s = sub (mod 4) in index order, the quad combines them
  // as (q0 + q1) + (q2 + q3) -- a fixed order -- and lane 0 of the quad does the Adam update.  With one thread per parameter
  // the kernel had ~16 KB of loads in flight per SM and read its 72 MB at 1.5 TB/s.
  const long long n_lanes = 4LL * kNsfParams;
  for (long long T = blockIdx.x * (long long)blockDim.x + threadIdx.x; T < ((n_lanes + 31) / 32) * 32;
       T += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(T >> 2), sub = (int)(T & 3);
    // locate the parameter and gather its (scaled) gradient
    float g = 0.f;
    int layer = -1, r = 0, c = 0;
    bool is_w = false, skip = t >= kNsfParams;
    auto strided_sum = [&](const float* q, size_t stride) {
      float acc = 0.f;
      for (int s = sub; s < a.splits; s += 4) acc += q[(size_t)s * stride];
      return acc;
    };
    auto small_sum = [&](int l, int m, int comp) {
      return strided_sum(a.small_part + ((size_t)l * a.splits * 128 + m) * 4 + comp, 512);
    };
    if (skip) {}
    else if (t < 384) { g = small_sum(0, t / 3, t % 3); }                  // dW0[m][c]
    else if (t < 512) { g = small_sum(0, t - 384, 3); }                    // db0
    else if (t >= nsf_off_b(8)) { skip = true; }                           // db8: the first warp of block 0, below
    else if (t >= nsf_off_w(8)) {
      const int k = t - nsf_off_w(8);                                      // dW8[k / 128][k % 128] = (h_8^T dflow)[j][k]
      g = small_sum(8, k & 127, k >> 7);
    } else {
      const int u = t - 512;
      layer = 1 + u / (128 * 128 + 128);
      const int w = u % (128 * 128 + 128);
      if (w < 128 * 128) {
        is_w = true; r = w / 128; c = w % 128;
        g = strided_sum(a.dW_part + ((size_t)(layer - 1) * a.splits) * 16384 + w, 16384);
      } else {
        g = small_sum(layer, w - 128 * 128, 3);                            // db_l
      }
    }
    const float g1 = g + __shfl_xor_sync(0xffffffffu, g, 1);               // (q0 + q1) in lanes 0,1; (q2 + q3) in lanes 2,3
    g = g1 + __shfl_xor_sync(0xffffffffu, g1, 2);
    if (skip || sub != 0) continue;
    g *= a.inv_scale;
    const float m = a.beta1 * a.m[t] + (1.f - a.beta1) * g;          // exp_avg.lerp_(grad, 1-beta1)
    const float v = a.beta2 * a.v[t] + (1.f - a.beta2) * g * g;      // exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2)
    a.m[t] = m; a.v[t] = v;
    const float denom = sqrtf(v) / sqrtf(bc2) + a.eps;
    const float pnew = a.params[t] - (a.lr / bc1) * (m / denom);
    a.params[t] = pnew;
    if (is_w) {
      const float scaled = a.planes == 2 ? pnew * kNsfWScale : pnew;
      umma::store_split(a.Wp[layer] + r * 128 + c, 16384, a.planes, scaled);
      umma::store_split(a.WpT[layer] + c * 128 + r, 16384, a.planes, scaled);
    }
  }
  // db8 = sum over the head blocks of their partials: 32 strided sums + a fixed shuffle tree (a single thread walking the
  // ~800 blocks was the longest chain of the kernel)
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    for (int k = 0; k < 3; ++k) {
      float g = 0.f;
      for (int blk = threadIdx.x; blk < a.head_blocks; blk += 32) g += a.head_part[(size_t)blk * kHeadPart + k];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) g += __shfl_xor_sync(0xffffffffu, g, d);
      if (threadIdx.x == 0) {
        const int t = nsf_off_b(8) + k;
        g *= a.inv_scale;
        const float m = a.beta1 * a.m[t] + (1.f - a.beta1) * g;
        const float v = a.beta2 * a.v[t] + (1.f - a.beta2) * g * g;
        a.m[t] = m; a.v[t] = v;
        a.params[t] -= (a.lr / bc1) * (m / (sqrtf(v) / sqrtf(bc2) + a.eps));
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_nsf_pack_weights(const float* __restrict__ params, NsfAdamArgs a) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 7 * 16384; t += gridDim.x * blockDim.x) {
    const int layer = 1 + t / 16384, w = t % 16384, r = w / 128, c = w % 128;
    const float pv = params[nsf_off_w(layer) + w];
    const float scaled = a.planes == 2 ? pv * kNsfWScale : pv;
    umma::store_split(a.Wp[layer] + r * 128 + c, 16384, a.planes, scaled);
    umma::store_split(a.WpT[layer] + c * 128 + r, 16384, a.planes, scaled);
  }
}

__global__ void __launch_bounds__(256)
k_nsf_pack_points(const float* __restrict__ pc, int n, int n_pad, float4* __restrict__ x4, __nv_bfloat16* __restrict__ x16,
                  int planes) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
    const bool live = i < n;
    const float4 p = live ? make_float4(pc[3 * (size_t)i], pc[3 * (size_t)i + 1], pc[3 * (size_t)i + 2], 0.f)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    x4[i] = p;
    // (x, y, z, 1, 0 x 28) as split planes: the B operand of k_nsf_dw's 16-column accumulator (dW_0 and every bias gradient)
    uint32_t hi[2], lo[2];
    umma::pack_split2(p.x, p.y, planes == 2, hi[0], lo[0]);
    umma::pack_split2(p.z, live ? 1.f : 0.f, planes == 2, hi[1], lo[1]);
    uint4* d0 = (uint4*)(x16 + (size_t)i * 32);
    d0[0] = make_uint4(hi[0], hi[1], 0u, 0u); d0[1] = make_uint4(0u, 0u, 0u, 0u);
    d0[2] = make_uint4(0u, 0u, 0u, 0u); d0[3] = make_uint4(0u, 0u, 0u, 0u);
    if (planes == 2) {
      uint4* d1 = (uint4*)(x16 + (size_t)n_pad * 32 + (size_t)i * 32);
      d1[0] = make_uint4(lo[0], lo[1], 0u, 0u); d1[1] = make_uint4(0u, 0u, 0u, 0u);
      d1[2] = make_uint4(0u, 0u, 0u, 0u); d1[3] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

__global__ void __launch_bounds__(256)
k_nsf_unpack_flow(const float* __restrict__ f4, int n, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    out[3 * (size_t)i] = f4[4 * (size_t)i]; out[3 * (size_t)i + 1] = f4[4 * (size_t)i + 1];
    out[3 * (size_t)i + 2] = f4[4 * (size_t)i + 2];
  }
}

struct NsfLayout {
  NsfBufs b;
  NsfAdamArgs ad;
  float* dW_part; int splits; int k_split;
  unsigned* bbox; NsfVol* vol_dev;
  int head_blocks;
};

static size_t nsf_layout(int n_max, int planes, NsfLayout* L, void* base) {
  Arena A(base, (size_t)-1);
  const int n_pad = ceil_div(n_max > 0 ? n_max : 1, 256) * 256;   // k_mlp_chain works on 256-point CTA-pair tiles
  const size_t act = (size_t)planes * n_pad * 128;
  NsfLayout l;
  l.b.x4 = A.take<float4>(n_pad);
  l.b.x16 = A.take<__nv_bfloat16>((size_t)planes * n_pad * 32);
  l.b.d16 = A.take<__nv_bfloat16>((size_t)planes * n_pad * 32);
  for (int k = 1; k <= kNsfLayers; ++k) { l.b.H[k] = A.take<__nv_bfloat16>(act); l.b.DL[k] = A.take<__nv_bfloat16>(act); }
  l.b.H[0] = l.b.DL[0] = nullptr;
  l.b.relu_mask[0] = nullptr;
  for (int k = 1; k < kNsfLayers; ++k) l.b.relu_mask[k] = A.take<uint32_t>((size_t)n_pad * 4);
  l.b.params = A.take<float>(kNsfParams);
  l.ad.m = A.take<float>(kNsfParams);
  l.ad.v = A.take<float>(kNsfParams);
  l.b.flow = A.take<float>((size_t)n_pad * 4);
  l.b.best_flow = A.take<float>((size_t)n_pad * 4);
  l.head_blocks = ceil_div(n_pad, kHeadThreads);
  l.b.head_part = A.take<float>((size_t)l.head_blocks * kHeadPart);
  l.b.gW0 = A.take<float>(128 * 3 + 128);
  l.b.gb = A.take<float>(7 * 128);
  l.b.ctl = A.take<NsfCtl>(1);
  l.k_split = 32 * ceil_div(n_pad, 32 * kNumSMs);
  l.splits = ceil_div(n_pad, l.k_split);
  l.dW_part = A.take<float>((size_t)7 * l.splits * 16384);
  l.b.small_part = A.take<float>((size_t)9 * l.splits * 128 * 4);
  for (int k = 1; k < kNsfLayers; ++k) {
    l.ad.Wp[k] = A.take<__nv_bfloat16>((size_t)planes * 16384);
    l.ad.WpT[k] = A.take<__nv_bfloat16>((size_t)planes * 16384);
  }
  l.ad.Wp[0] = l.ad.WpT[0] = nullptr;
  l.bbox = A.take<unsigned>(8);
  l.vol_dev = A.take<NsfVol>(1);
  if (L) *L = l;
  return A.off + 1024;
}

#define HIMO_RET(expr) do { int _s = (expr); if (_s != HIMO_OK) return _s; } while (0)

static int nsf_gemm(const himo_conv_desc& d, cudaStream_t stream) { return himo_conv2d_nhwc(&d, stream); }

static int g_nsf_fused = 1;
static PFN_cuTensorMapEncodeTiled nsf_encode_fn();
// [planes][rows][C] 16-bit tensor, box = (32 channels, box_rows, 1 plane), 64-byte swizzle
static int nsf_map(CUtensorMap* m, const void* base, int planes, long long rows, int C, int box_rows) {
  PFN_cuTensorMapEncodeTiled enc = nsf_encode_fn();
  if (!enc) return HIMO_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)rows * C * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? HIMO_OK : HIMO_ERR_ARG;
}

template <int DIR>
static int nsf_launch_chain(const NsfBufs& b, const NsfAdamArgs& ad, cudaStream_t stream) {
  ChainMaps maps;
  for (int l = 1; l < kNsfLayers; ++l) HIMO_RET(nsf_map(&maps.w[l], DIR == 0 ? ad.Wp[l] : ad.WpT[l], 2, 128, 128, 64));
  maps.w[0] = maps.w[1];
  for (int l = 1; l <= kNsfLayers; ++l) HIMO_RET(nsf_map(&maps.act[l], DIR == 0 ? b.H[l] : b.DL[l], 2, b.n_pad, 128, 128));
  maps.act[0] = maps.act[1];
  ChainParams p;
  p.n_pair_tiles = b.n_pad / 256;
  p.params = b.params; p.x4 = b.x4; p.acc_scale = 1.0f / kNsfWScale; p.stop_flag = &b.ctl->stop;
  for (int l = 0; l < kNsfLayers; ++l) p.mask[l] = b.relu_mask[l];
  static bool configured_dev[64] = {};
  int dev_ = 0;
  HIMO_CUDA_RET(cudaGetDevice(&dev_));
  if (!configured_dev[dev_ & 63]) {
    HIMO_CUDA_RET(cudaFuncSetAttribute(k_mlp_chain<DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChTotal));
    configured_dev[dev_ & 63] = true;
  }
  const int pairs = p.n_pair_tiles < kNumSMs / 2 ? p.n_pair_tiles : kNumSMs / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pairs * 2); cfg.blockDim = dim3(kChThreads); cfg.dynamicSmemBytes = kChTotal; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, k_mlp_chain<DIR>, maps, p));
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

// h_1 = relu(W0 x + b0), then h_{l+1} = relu(W_l h_l + b_l), l = 1..7: one fused chain launch in the split-plane mode
// (k_mlp_chain<0>), else layer 0 on CUDA cores + seven launches of the tcgen05 GEMM
static int nsf_forward_hidden(const NsfBufs& b, const NsfAdamArgs& ad, cudaStream_t stream) {
  const int P = b.planes, n_pad = b.n_pad;
  if (P == 2 && g_nsf_fused) return nsf_launch_chain<0>(b, ad, stream);
  const float wscale = P == 2 ? 1.0f / kNsfWScale : 1.0f;
  k_nsf_l0_fwd<<<min(n_pad / kL0Pts, kNumSMs * 8), 256, 0, stream>>>(b); HIMO_LAUNCH_RET();
  for (int l = 1; l < kNsfLayers; ++l) {
    himo_conv_desc g = {};
    g.in = b.H[l]; g.in_planes = P; g.in_plane_stride = b.ps; g.H_in = n_pad / 128; g.W_in = 128;
    g.Cin_total = 128; g.Cin = 128; g.wgt = ad.Wp[l]; g.bias = b.params + nsf_off_b(l); g.Cout = 128; g.ksize = 1;
    g.stride = 1; g.out = b.H[l + 1]; g.out_planes = P; g.out_plane_stride = b.ps; g.Cout_total = 128; g.act = 4;
    g.n_groups = 1; g.acc_scale = wscale; g.stop_flag = &b.ctl->stop;
    HIMO_RET(nsf_gemm(g, stream));
  }
  return HIMO_OK;
}

static PFN_cuTensorMapEncodeTiled nsf_encode_fn() {
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
// [planes][n_pad][C] 16-bit tensor, box = (32 features, kDwKT points, 1 plane), 64-byte swizzle
static int nsf_map_rows(CUtensorMap* m, const void* base, int planes, int n_pad, int C) {
  PFN_cuTensorMapEncodeTiled enc = nsf_encode_fn();
  if (!enc) return HIMO_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n_pad, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)n_pad * C * 2};
  cuuint32_t box[3] = {32, (cuuint32_t)kDwKT, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? HIMO_OK : HIMO_ERR_ARG;
}

// From delta_8 in DL[8]: delta_l = (delta_{l+1} W_l) * relu'(h_l) for l = 7..1 (GEMMs with the ReLU mask in the epilogue),
// then every weight / bias gradient of layers 0..7 in one launch (k_nsf_dw) and the small reductions.
static int nsf_backward_hidden(const NsfBufs& b, const NsfAdamArgs& ad, float* dW_part, int splits, int k_split,
                               cudaStream_t stream) {
  const int P = b.planes, n_pad = b.n_pad;
  const float wscale = P == 2 ? 1.0f / kNsfWScale : 1.0f;
  if (P == 2 && g_nsf_fused) HIMO_RET(nsf_launch_chain<1>(b, ad, stream));
  else for (int l = kNsfLayers - 1; l >= 1; --l) {
    himo_conv_desc q = {};
    q.in = b.DL[l + 1]; q.in_planes = P; q.in_plane_stride = b.ps; q.H_in = n_pad / 128; q.W_in = 128;
    q.Cin_total = 128; q.Cin = 128; q.wgt = ad.WpT[l]; q.bias = nullptr; q.Cout = 128; q.ksize = 1; q.stride = 1;
    q.out = b.DL[l]; q.out_planes = P; q.out_plane_stride = b.ps; q.Cout_total = 128; q.act = 0;
    q.n_groups = 1; q.acc_scale = wscale;
    q.mask_src = b.H[l]; q.mask_plane_stride = b.ps; q.mask_planes = P; q.stop_flag = &b.ctl->stop;
    HIMO_RET(nsf_gemm(q, stream));
  }
  DwMaps maps;
  for (int l = 1; l <= kNsfLayers; ++l) HIMO_RET(nsf_map_rows(&maps.delta[l], b.DL[l], P, n_pad, 128));
  maps.delta[0] = maps.delta[1];
  for (int l = 1; l <= kNsfLayers; ++l) HIMO_RET(nsf_map_rows(&maps.h[l], b.H[l], P, n_pad, 128));
  maps.h[0] = maps.h[1];
  HIMO_RET(nsf_map_rows(&maps.x, b.x16, P, n_pad, 32));
  HIMO_RET(nsf_map_rows(&maps.d, b.d16, P, n_pad, 32));
  static bool configured_dev[64] = {};
  int dev_ = 0;
  HIMO_CUDA_RET(cudaGetDevice(&dev_));
  if (!configured_dev[dev_ & 63]) {
    HIMO_CUDA_RET(cudaFuncSetAttribute(k_nsf_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwTotal));
    configured_dev[dev_ & 63] = true;
  }
  k_nsf_dw<<<splits, kDwThreads, kDwTotal, stream>>>(maps, &b.ctl->stop, n_pad, k_split, splits, P, dW_part, b.small_part);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

// Per-call view of a workspace: buffers, Adam arguments, split-K geometry for `n` points.
struct NsfCall { NsfLayout L; NsfBufs b; NsfAdamArgs ad; int head_blocks, k_split, splits; };
static int nsf_call_setup(void* workspace, size_t workspace_bytes, int n_max, int planes, int n, float lr, NsfCall* c) {
  if (!workspace || (planes != 1 && planes != 2) || n <= 0 || n > n_max) return HIMO_ERR_ARG;
  if (nsf_layout(n_max, planes, &c->L, workspace) > workspace_bytes) return HIMO_ERR_WORKSPACE;
  const int n_pad = ceil_div(n, 256) * 256;
  c->b = c->L.b;
  c->b.n = n; c->b.n_pad = n_pad; c->b.planes = planes; c->b.ps = (long long)n_pad * 128;
  int e = 0; while ((1 << e) < n) ++e;
  c->b.grad_scale = (float)(1 << e);
  c->head_blocks = ceil_div(n_pad, kHeadThreads);
  c->k_split = 32 * ceil_div(n_pad, 32 * kNumSMs);
  c->splits = ceil_div(n_pad, c->k_split);
  c->ad = c->L.ad;
  c->ad.params = c->b.params; c->ad.dW_part = c->L.dW_part; c->ad.splits = c->splits; c->ad.gb = c->b.gb;
  c->ad.gW0 = c->b.gW0; c->ad.gb0 = c->b.gW0 + 384; c->ad.small_part = c->b.small_part; c->ad.head_part = c->b.head_part; c->ad.head_blocks = c->head_blocks;
  c->ad.planes = planes; c->ad.lr = lr; c->ad.beta1 = 0.9f; c->ad.beta2 = 0.999f; c->ad.eps = 1e-8f;
  c->ad.inv_scale = 1.0f / c->b.grad_scale; c->ad.ctl = c->b.ctl;
  return HIMO_OK;
}

}  // namespace himo

using namespace himo;

// A/B knob: 0 runs the hidden layers as 14 GEMM launches per iteration instead of the two fused chain kernels.
extern "C" int himo_nsf_set_fused(int enable) { g_nsf_fused = enable ? 1 : 0; return HIMO_OK; }

extern "C" size_t himo_nsf_workspace_bytes(int n_max, int planes) {
  if (n_max < 0 || (planes != 1 && planes != 2)) return 0;
  return nsf_layout(n_max, planes, nullptr, nullptr);
}

// Distance-volume geometry of a frame pair: host gets lo[3], dims[3] (one small synchronous copy per
// frame pair; the reference syncs >= 2x per ITERATION).
extern "C" int himo_nsf_volume_geometry(const float* pc0, int n0, const float* pc1, int n1, float grid_factor,
                                        float* lo_host, int32_t* dims_host, void* workspace, void* stream_) {
  if (n0 <= 0 || n1 <= 0 || !pc0 || !pc1 || !workspace || !lo_host || !dims_host) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  unsigned* bbox = (unsigned*)workspace;
  NsfVol* vd = (NsfVol*)((char*)workspace + 256);
  k_nsf_bbox_init<<<1, 32, 0, stream>>>(bbox); HIMO_LAUNCH_RET();
  k_nsf_bbox<<<min(ceil_div(n0 + n1, 256), kNumSMs * 4), 256, 0, stream>>>(pc0, n0, pc1, n1, bbox); HIMO_LAUNCH_RET();
  k_nsf_vol<<<1, 32, 0, stream>>>(bbox, grid_factor, vd); HIMO_LAUNCH_RET();
  NsfVol hv;
  HIMO_CUDA_RET(cudaMemcpyAsync(&hv, vd, sizeof(NsfVol), cudaMemcpyDeviceToHost, stream));
  HIMO_CUDA_RET(cudaStreamSynchronize(stream));
  for (int k = 0; k < 3; ++k) { lo_host[k] = hv.lo[k]; dims_host[k] = hv.dims[k]; }
  return HIMO_OK;
}

static int g_nsf_blocking_poll = 0;
extern "C" int himo_nsf_set_blocking_poll(int enable) { g_nsf_blocking_poll = enable ? 1 : 0; return HIMO_OK; }
static int g_nsf_head_warp = 1;
extern "C" int himo_nsf_set_head_warp(int enable) { g_nsf_head_warp = enable ? 1 : 0; return HIMO_OK; }
static int g_dt_cluster = 1;
static int g_dt_big_tiles = 3;   // variant of the tiled pass on planes of >= 256 k cells; per axis-2 direction at 1040 x 1030 x 52:
                                 // 0 = 16x16 tiles x 16 planes 1.90 ms, 1 = 32x32 x 16: 3.87, 2 = 16x16 x 8: 1.41, 3 = 16x16 x 4: 1.28
extern "C" int himo_nsf_set_dt_big_tiles(int variant) { g_dt_big_tiles = variant; return HIMO_OK; }
static long long* g_dt_dbg = nullptr;   // device buffer [16 CTAs][4 probes][4] or null
extern "C" int himo_nsf_set_dt_debug_buffer(long long* p) { g_dt_dbg = p; return HIMO_OK; }
extern "C" int himo_nsf_set_dt_cluster(int enable) { g_dt_cluster = enable ? 1 : 0; return HIMO_OK; }

// Row stride of the shared-memory state: a warp's 32 strips cover the tail of one 4-row group and the head of the next; with
// 4 * SW == nw (mod 32) the head continues the bank sequence of the tail instead of colliding with it (ncu: 52 % of the
// shared loads conflicted with SW = nw + 2 = 54).
static int dt_sweep_row_stride(int nw) {
  int sw = nw + 2;
  if (nw % 4 == 0) while ((4 * sw - nw) % 32 != 0) ++sw;
  return sw;
}

// One (axis, direction) pass as a single cluster launch; HIMO_ERR_UNSUPPORTED when the plane does not fit a cluster.
static int dt_sweep_launch(float* D, const int* n, int axis, int dir, float l00, float l01, float l11, cudaStream_t stream) {
  const int nh = axis == 0 ? n[1] : n[0], nw = n[2];
  static int max_cluster[64] = {};                            // per device: 16 where allowed, else 8, -1 = none
  int dev = 0;
  HIMO_CUDA_RET(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return HIMO_ERR_UNSUPPORTED;
  for (int C = max_cluster[dev] > 0 ? max_cluster[dev] : 16; C >= 8; C >>= 1) {
    if (max_cluster[dev] < 0) break;
    const int R = ceil_div(nh, C);
    if ((long long)ceil_div(R, kDtcVS) * nw > (long long)kDtcMaxStrips * kDtcThreads) continue;
    const int SW = dt_sweep_row_stride(nw);
    const size_t smem = (3 * (size_t)(R + 2) * SW + (size_t)kDtcPF * R * nw) * sizeof(float);
    if (smem > 200 * 1024) continue;
    if (max_cluster[dev] == 0) {                              // first call on this device: what cluster size may launch?
      HIMO_CUDA_RET(cudaFuncSetAttribute(k_nsf_dt_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (C > 8 && cudaFuncSetAttribute(k_nsf_dt_sweep, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(C); cfg.blockDim = dim3(kDtcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (max_cluster[dev] == 0) {
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, k_nsf_dt_sweep, &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        continue;
      }
      max_cluster[dev] = C;
    }
    HIMO_CUDA_RET(cudaLaunchKernelEx(&cfg, k_nsf_dt_sweep, D, n[0], n[1], n[2], axis, dir, l00, l01, l11, R, C, SW, g_dt_dbg));
    himo_count_launch_();
    return HIMO_OK;
  }
  if (max_cluster[dev] == 0) max_cluster[dev] = -1;
  return HIMO_ERR_UNSUPPORTED;
}

// One (axis, direction) pass as tiled launches of 16 planes each; on big planes (the axis-2 pass: in-plane strides of 52 and
// 53 560 floats, nothing coalesces) fewer planes per launch win: the halo cone and with it the recomputation shrink.
static int dt_tiled_pass(float* D, const int* n, int axis, int dir, float l00, float l01, float l11, cudaStream_t stream) {
  const int h = axis == 0 ? 1 : 0, w = axis == 2 ? 1 : 2;
  // variant 0: 16x16 tiles, 16 planes per launch; 1: 32x32 tiles, 16 planes; 2: 16x16 tiles, 8 planes (halo 8: 4x recomputation)
  const int variant = (long long)n[h] * n[w] >= (1 << 18) ? g_dt_big_tiles : 0;
  const int tile = variant == 1 ? kDtTileBig : kDtTile;
  const int steps = variant == 2 ? 8 : (variant == 3 ? 4 : kDtStepsDefault);
  dim3 grid(ceil_div(n[w], tile), ceil_div(n[h], tile));
  int p = dir > 0 ? 1 : n[axis] - 2;
  int remaining = n[axis] - 1;
  while (remaining > 0) {
    const int cnt = remaining < steps ? remaining : steps;
    if (variant == 1) k_nsf_dt_pass<kDtTileBig, 16><<<grid, 256, 0, stream>>>(D, n[0], n[1], n[2], axis, dir, p, cnt, l00, l01, l11);
    else if (variant == 2) k_nsf_dt_pass<kDtTile, 8><<<grid, 256, 0, stream>>>(D, n[0], n[1], n[2], axis, dir, p, cnt, l00, l01, l11);
    else if (variant == 3) k_nsf_dt_pass<kDtTile, 4><<<grid, 256, 0, stream>>>(D, n[0], n[1], n[2], axis, dir, p, cnt, l00, l01, l11);
    else k_nsf_dt_pass<kDtTile, 16><<<grid, 256, 0, stream>>>(D, n[0], n[1], n[2], axis, dir, p, cnt, l00, l01, l11);
    HIMO_LAUNCH_RET();
    p += dir * cnt;
    remaining -= cnt;
  }
  return HIMO_OK;
}

// One raster pass in place, either way (tests / debugging): sweep = 1 -> k_nsf_dt_sweep, 0 -> the tiled launches.
extern "C" int himo_nsf_dt_pass(float* D, const int32_t* dims, float grid_factor, int axis, int dir, int sweep, void* stream_) {
  if (!D || !dims || axis < 0 || axis > 2 || (dir != 1 && dir != -1)) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const float sp = 1.0f / grid_factor;
  const float l00 = sqrtf(sp * sp), l01 = sqrtf(sp * sp + sp * sp), l11 = sqrtf(sp * sp + sp * sp + sp * sp);
  const int n[3] = {dims[0], dims[1], dims[2]};
  if (n[axis] < 2) return HIMO_OK;
  if (sweep) return axis < 2 ? dt_sweep_launch(D, n, axis, dir, l00, l01, l11, stream) : HIMO_ERR_UNSUPPORTED;
  return dt_tiled_pass(D, n, axis, dir, l00, l01, l11, stream);
}

// D[H][W][D] = FastGeodis-style raster Euclidean distance transform of the occupancy of pc1.
extern "C" int himo_nsf_dt_build(const float* pc1, int n1, const float* lo, const int32_t* dims, float grid_factor,
                                 float* D, void* stream_) {
  if (!pc1 || n1 < 0 || !lo || !dims || !D) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfVol v;
  for (int k = 0; k < 3; ++k) { v.lo[k] = lo[k]; v.dims[k] = dims[k]; if (dims[k] <= 0) return HIMO_ERR_ARG; }
  v.gf = grid_factor;
  const long long total = (long long)dims[0] * dims[1] * dims[2];
  k_nsf_fill<<<kNumSMs * 8, 256, 0, stream>>>(D, total, 1e10f); HIMO_LAUNCH_RET();
  if (n1 > 0) { k_nsf_occupancy<<<min(ceil_div(n1, 256), kNumSMs * 8), 256, 0, stream>>>(pc1, n1, v, D); HIMO_LAUNCH_RET(); }
  const float sp = 1.0f / grid_factor;
  const float l00 = sqrtf(sp * sp), l01 = sqrtf(sp * sp + sp * sp), l11 = sqrtf(sp * sp + sp * sp + sp * sp);
  const int n[3] = {dims[0], dims[1], dims[2]};
  for (int axis = 0; axis < 3; ++axis) {
    for (int dir = 1; dir >= -1; dir -= 2) {
      if (axis < 2 && n[axis] > 1 && g_dt_cluster) {          // small planes: one cluster sweeps the whole pass
        const int st = dt_sweep_launch(D, n, axis, dir, l00, l01, l11, stream);
        if (st == HIMO_OK) continue;
        if (st != HIMO_ERR_UNSUPPORTED) return st;            // does not fit (huge planes): the tiled passes below
      }
      HIMO_RET(dt_tiled_pass(D, n, axis, dir, l00, l01, l11, stream));
    }
  }
  return HIMO_OK;
}

// The optimisation loop.  Blocking call (it polls the device stop flag every `poll` iterations).
extern "C" int himo_nsf_optimize(const himo_nsf_desc* d, void* stream_) {
  if (!d || !d->pc0 || d->n <= 0 || !d->D || !d->init_params || !d->best_flow || !d->workspace) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int P = d->planes;
  if (P != 1 && P != 2) return HIMO_ERR_ARG;
  NsfLayout L;
  if (nsf_layout(d->n_max, P, &L, d->workspace) > d->workspace_bytes || d->n > d->n_max) return HIMO_ERR_WORKSPACE;
  const int n = d->n, n_pad = ceil_div(n, 256) * 256;
  NsfBufs b = L.b;
  b.n = n; b.n_pad = n_pad; b.planes = P; b.ps = (long long)n_pad * 128;
  int e = 0; while ((1 << e) < n) ++e;
  b.grad_scale = (float)(1 << e);
  NsfVol vol;
  for (int k = 0; k < 3; ++k) { vol.lo[k] = d->lo[k]; vol.dims[k] = d->dims[k]; }
  vol.gf = d->grid_factor;
  const int head_blocks = ceil_div(n_pad, kHeadThreads);
  const int k_split = 32 * ceil_div(n_pad, 32 * kNumSMs);
  const int splits = ceil_div(n_pad, k_split);

  NsfAdamArgs ad = L.ad;
  ad.params = b.params; ad.dW_part = L.dW_part; ad.splits = splits; ad.gb = b.gb; ad.gW0 = b.gW0; ad.gb0 = b.gW0 + 384; ad.small_part = b.small_part;
  ad.head_part = b.head_part; ad.head_blocks = head_blocks; ad.planes = P;
  ad.lr = d->lr; ad.beta1 = 0.9f; ad.beta2 = 0.999f; ad.eps = 1e-8f; ad.inv_scale = 1.0f / b.grad_scale; ad.ctl = b.ctl;

  // ---- state init
  HIMO_CUDA_RET(cudaMemcpyAsync(b.params, d->init_params, sizeof(float) * kNsfParams, cudaMemcpyDeviceToDevice, stream));
  HIMO_CUDA_RET(cudaMemsetAsync(ad.m, 0, sizeof(float) * kNsfParams, stream));
  HIMO_CUDA_RET(cudaMemsetAsync(ad.v, 0, sizeof(float) * kNsfParams, stream));
  NsfCtl c0 = {};
  c0.best_loss = INFINITY;
  HIMO_CUDA_RET(cudaMemcpyAsync(b.ctl, &c0, sizeof(NsfCtl), cudaMemcpyHostToDevice, stream));
  HIMO_CUDA_RET(cudaMemsetAsync(b.best_flow, 0, sizeof(float) * 4 * n_pad, stream));
  k_nsf_pack_points<<<min(ceil_div(n_pad, 256), kNumSMs * 8), 256, 0, stream>>>(d->pc0, n, n_pad, (float4*)b.x4, b.x16, P);
  HIMO_LAUNCH_RET();
  k_nsf_pack_weights<<<kNumSMs * 2, 256, 0, stream>>>(b.params, ad);
  HIMO_LAUNCH_RET();

  auto iteration = [&]() -> int {
    HIMO_RET(nsf_forward_hidden(b, ad, stream));
    if (g_nsf_head_warp) k_nsf_head_warp<<<head_blocks, kHeadThreads, 0, stream>>>(b, d->D, vol);
    else k_nsf_head<<<head_blocks, kHeadThreads, 0, stream>>>(b, d->D, vol);
    HIMO_LAUNCH_RET();
    k_nsf_control<<<1, 32, 0, stream>>>(b, head_blocks, d->min_delta, d->patience); HIMO_LAUNCH_RET();
    k_nsf_snapshot<<<min(ceil_div(n, 256), kNumSMs * 4), 256, 0, stream>>>(b); HIMO_LAUNCH_RET();
    HIMO_RET(nsf_backward_hidden(b, ad, L.dW_part, splits, k_split, stream));
    k_nsf_adam<<<kNumSMs * 8, 256, 0, stream>>>(ad); HIMO_LAUNCH_RET();
    return HIMO_OK;
  };

  NsfCtl hc = {};
  const int poll = d->poll_iters > 0 ? d->poll_iters : 8;
  int launched = 0;
  // The host waits for every chunk on a BLOCKING-sync event (the thread sleeps instead of spinning in
  // cudaStreamSynchronize): several pairs are optimised from worker threads of one process, one process per GPU, and
  // spinning pollers would take a host core each for the whole run.
  // (himo_nsf_set_blocking_poll(1), set by FastNSFEngine.infer_stream; a lone optimiser keeps the spinning wait: the sleeping
  //  one costs ~0.05 ms per iteration in wake-up latency, 0.49 -> 0.54 ms)
  cudaEvent_t chunk_done = nullptr;
  HIMO_CUDA_RET(cudaEventCreateWithFlags(&chunk_done, (g_nsf_blocking_poll ? cudaEventBlockingSync : 0) | cudaEventDisableTiming));
  int loop_status = HIMO_OK;
  while (launched < d->max_iters) {
    const int chunk = (d->max_iters - launched) < poll ? (d->max_iters - launched) : poll;
    for (int k = 0; k < chunk && loop_status == HIMO_OK; ++k) loop_status = iteration();
    if (loop_status != HIMO_OK) break;
    launched += chunk;
    cudaError_t ce = cudaEventRecord(chunk_done, stream);
    if (ce == cudaSuccess) ce = cudaEventSynchronize(chunk_done);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&hc, b.ctl, sizeof(NsfCtl), cudaMemcpyDeviceToHost, stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
    if (ce != cudaSuccess) { loop_status = (int)ce; break; }     // > 0: cudaError_t, as everywhere in this ABI
    if (hc.stop) break;
  }
  cudaEventDestroy(chunk_done);
  if (loop_status != HIMO_OK) return loop_status;
  k_nsf_unpack_flow<<<min(ceil_div(n, 256), kNumSMs * 4), 256, 0, stream>>>(b.best_flow, n, d->best_flow);
  HIMO_LAUNCH_RET();
  if (d->final_params)
    HIMO_CUDA_RET(cudaMemcpyAsync(d->final_params, b.params, sizeof(float) * kNsfParams, cudaMemcpyDeviceToDevice, stream));
  if (d->exp_avg_out)
    HIMO_CUDA_RET(cudaMemcpyAsync(d->exp_avg_out, ad.m, sizeof(float) * kNsfParams, cudaMemcpyDeviceToDevice, stream));
  HIMO_CUDA_RET(cudaStreamSynchronize(stream));
  if (d->iterations_out) *d->iterations_out = hc.iters;
  if (d->best_loss_out) *d->best_loss_out = hc.best_loss;
  if (d->last_loss_out) *d->last_loss_out = hc.loss;
  return HIMO_OK;
}

// ---------------------------------------------------------------------------------------------- generic MLP C ABI
// replaces: Neural_Prior.forward + autograd + torch.optim.Adam as NSFP uses them (OSF/src/models/nsfp.py:74-131,
// basic/nsfp_module.py:7-47).  All state (parameters, Adam moments, activations, control block) lives in the
// caller's workspace (himo_nsf_workspace_bytes); `ctl_workspace` (NULL = own) lets a second network follow the
// control block -- iteration count and stop flag -- of the first (NSFP trains net and net_inv in lockstep).
static NsfCtl* mlp_ctl(void* ctl_ws, int n_max, int planes, NsfCtl* own) {
  if (!ctl_ws) return own;
  NsfLayout o;
  nsf_layout(n_max, planes, &o, ctl_ws);
  return o.b.ctl;
}

extern "C" int himo_mlp_init(void* workspace, size_t workspace_bytes, int n_max, int planes, const float* init_params,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfCall c;
  if (!init_params) return HIMO_ERR_ARG;
  HIMO_RET(nsf_call_setup(workspace, workspace_bytes, n_max, planes, 1, 0.f, &c));
  HIMO_CUDA_RET(cudaMemcpyAsync(c.b.params, init_params, sizeof(float) * kNsfParams, cudaMemcpyDeviceToDevice, stream));
  HIMO_CUDA_RET(cudaMemsetAsync(c.ad.m, 0, sizeof(float) * kNsfParams, stream));
  HIMO_CUDA_RET(cudaMemsetAsync(c.ad.v, 0, sizeof(float) * kNsfParams, stream));
  NsfCtl c0 = {};
  c0.best_loss = INFINITY;
  HIMO_CUDA_RET(cudaMemcpyAsync(c.b.ctl, &c0, sizeof(NsfCtl), cudaMemcpyHostToDevice, stream));
  k_nsf_pack_weights<<<kNumSMs * 2, 256, 0, stream>>>(c.b.params, c.ad);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

extern "C" int himo_mlp_forward(void* workspace, size_t workspace_bytes, int n_max, int planes, void* ctl_workspace,
                                const float* x, int n, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfCall c;
  if (!x || !out) return HIMO_ERR_ARG;
  HIMO_RET(nsf_call_setup(workspace, workspace_bytes, n_max, planes, n, 0.f, &c));
  c.b.ctl = mlp_ctl(ctl_workspace, n_max, planes, c.b.ctl);
  k_nsf_pack_points<<<min(ceil_div(c.b.n_pad, 256), kNumSMs * 8), 256, 0, stream>>>(x, n, c.b.n_pad, (float4*)c.b.x4,
                                                                                    c.b.x16, planes);
  HIMO_LAUNCH_RET();
  HIMO_RET(nsf_forward_hidden(c.b, c.ad, stream));
  k_mlp_head_fwd<<<c.head_blocks, kHeadThreads, 0, stream>>>(c.b, out); HIMO_LAUNCH_RET();
  return HIMO_OK;
}

extern "C" int himo_mlp_backward(void* workspace, size_t workspace_bytes, int n_max, int planes, void* ctl_workspace,
                                 int n, const float* d_out, float* d_x, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfCall c;
  if (!d_out) return HIMO_ERR_ARG;
  HIMO_RET(nsf_call_setup(workspace, workspace_bytes, n_max, planes, n, 0.f, &c));
  c.b.ctl = mlp_ctl(ctl_workspace, n_max, planes, c.b.ctl);
  k_mlp_head_bwd<<<c.head_blocks, kHeadThreads, 0, stream>>>(c.b, d_out); HIMO_LAUNCH_RET();
  HIMO_RET(nsf_backward_hidden(c.b, c.ad, c.L.dW_part, c.splits, c.k_split, stream));
  if (d_x) {
    k_mlp_dx<<<ceil_div(n, 128), 128, 0, stream>>>(c.b, c.ad.inv_scale, d_x);
    HIMO_LAUNCH_RET();
  }
  return HIMO_OK;
}

// torch.optim.Adam.step for this network with the gradients of the last himo_mlp_backward; the step number is the
// control block's iteration count (himo_mlp_control increments it once per loss evaluation, like the reference's loop).
extern "C" int himo_mlp_adam_step(void* workspace, size_t workspace_bytes, int n_max, int planes, void* ctl_workspace,
                                  int n, float lr, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfCall c;
  HIMO_RET(nsf_call_setup(workspace, workspace_bytes, n_max, planes, n, lr, &c));
  c.ad.ctl = mlp_ctl(ctl_workspace, n_max, planes, c.b.ctl);
  k_nsf_adam<<<kNumSMs * 8, 256, 0, stream>>>(c.ad); HIMO_LAUNCH_RET();
  return HIMO_OK;
}

// One loss evaluation of the optimisation loop: best-loss / early-stopping state machine on the device, then the
// best-output snapshot (`out` [count] floats -> `best_out`, both device) if this loss is the best so far.
extern "C" int himo_mlp_control(void* workspace, size_t workspace_bytes, int n_max, int planes, const float* loss_dev,
                                float min_delta, int patience, const float* out, float* best_out, long long count,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfCall c;
  if (!loss_dev) return HIMO_ERR_ARG;
  HIMO_RET(nsf_call_setup(workspace, workspace_bytes, n_max, planes, 1, 0.f, &c));
  k_mlp_control<<<1, 32, 0, stream>>>(c.b.ctl, loss_dev, min_delta, patience); HIMO_LAUNCH_RET();
  if (out && best_out && count > 0) {
    k_mlp_snapshot<<<(int)min((long long)kNumSMs * 4, ceil_div_ll(count, 256)), 256, 0, stream>>>(c.b.ctl, out, best_out, count);
    HIMO_LAUNCH_RET();
  }
  return HIMO_OK;
}

// Blocking read of the control block: state[0] = stop, [1] = iterations, [2] = best loss, [3] = last loss.
extern "C" int himo_mlp_read_state(void* workspace, size_t workspace_bytes, int n_max, int planes, float* state_host,
                                   float* params_out, float* exp_avg_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NsfCall c;
  if (!state_host) return HIMO_ERR_ARG;
  HIMO_RET(nsf_call_setup(workspace, workspace_bytes, n_max, planes, 1, 0.f, &c));
  NsfCtl hc;
  HIMO_CUDA_RET(cudaMemcpyAsync(&hc, c.b.ctl, sizeof(NsfCtl), cudaMemcpyDeviceToHost, stream));
  if (params_out) HIMO_CUDA_RET(cudaMemcpyAsync(params_out, c.b.params, sizeof(float) * kNsfParams, cudaMemcpyDeviceToDevice, stream));
  if (exp_avg_out) HIMO_CUDA_RET(cudaMemcpyAsync(exp_avg_out, c.ad.m, sizeof(float) * kNsfParams, cudaMemcpyDeviceToDevice, stream));
  HIMO_CUDA_RET(cudaStreamSynchronize(stream));
  state_host[0] = (float)hc.stop; state_host[1] = (float)hc.iters; state_host[2] = hc.best_loss; state_host[3] = hc.loss;
  return HIMO_OK;
}
