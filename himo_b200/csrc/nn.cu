// himo_b200/csrc/nn.cu -- H2: exact bidirectional 1-NN (Chamfer correspondence) on an implicit Morton octree.
//
// Drop-in, at the C ABI, for chamfer3D.forward / chamfer3D.backward
// (OSF/assets/cuda/chamfer3D/chamfer3D_cuda.cpp:18-35, chamfer3D.cu:33-154).
// The reference streams the whole other cloud past every query (O(N0*N1), 256-point smem tiles).
// Here both clouds are counting-sorted by the Morton key of a uniform cell grid: one occupancy bit per
// cell + popcount ranks give CSR cell ranges with no hashing and no radix sort, and because the key is a
// Morton code every aligned 2^k-cell cube (an octree node) is ONE contiguous key range, i.e. one contiguous
// run of sorted points, found with two rank lookups.  A query first looks at the 3x3x3 cells around it,
// then at the 3x3x3 nodes of every coarser level until the covered slab proves the best candidate global;
// nodes holding many points are descended depth-first, nearest child first, with box-distance pruning.
// Lidar clouds span 400 m with a few far returns: the grid covers a robust box (mean +- 4 sigma, clipped
// to the bounding box) and points outside are stored in the boundary cells -- projection onto a convex box
// is non-expansive, so every box bound computed from the clamped query stays a valid lower bound.
// The result is the *exact* nearest neighbour with the reference's tie rule (lowest index), and the
// squared distance is evaluated with the rounding sequence of the reference's compiled kernel,
// fma(dz,dz, fma(dx,dx, dy*dy)) (read off the SASS of oracle/_ref/chamfer3D.so: FMUL on dy, FFMA dx, FFMA dz).
#include "common.cuh"
#include "himo_b200.h"

namespace himo {

// key space: 2^26 cells (8 MiB bitmap per cloud) up to 400 k points in total, 2^28 beyond -- dense million-point
// clouds want 0.125 m cells (1M-point lidar pair: 5.8 ms at 0.25 m, 3.8 ms at 0.125 m)
__host__ __device__ inline int nn_max_bits(long long n_total) { return n_total < 400000 ? 26 : 28; }
__host__ __device__ inline long long nn_max_words(long long n_total) { return (1ll << nn_max_bits(n_total)) / 64 + 1; }
constexpr int kNNLeaf = 128;   // nodes with <= this many points are scanned: expanding a node costs ~50 point tests

struct NNGrid {
  float ox, oy, oz;   // grid origin
  float h;            // fine cell edge
  float eps;          // slack that absorbs fp32 rounding of the cell assignment
  int bxy, bz;        // bits per axis: x and y share bxy, bz <= bxy
  int n_words;        // 64-bit bitmap words in use
};

struct NNStats {      // accumulated over both clouds (finite points only)
  double sum[3], sumsq[3];
  unsigned long long count;
  unsigned lo[3], hi[3];   // order-preserving uint encodings of the min / max coordinates
};

// monotone float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_nn_stats_init(NNStats* st) {
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) { st->sum[k] = 0.0; st->sumsq[k] = 0.0; st->lo[k] = 0xffffffffu; st->hi[k] = 0u; }
    st->count = 0ull;
  }
}

__global__ void __launch_bounds__(256)
k_nn_stats(const float* __restrict__ pc0, int n0, const float* __restrict__ pc1, int n1, NNStats* __restrict__ st) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  double sm[3] = {0.0, 0.0, 0.0}, sq[3] = {0.0, 0.0, 0.0};
  unsigned cnt = 0;
  const int n = n0 + n1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = i < n0 ? pc0 + 3 * (size_t)i : pc1 + 3 * (size_t)(i - n0);
    const float v[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    if (!(isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]))) continue;
    ++cnt;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], v[k]); hi[k] = fmaxf(hi[k], v[k]);
      sm[k] += (double)v[k]; sq[k] += (double)v[k] * (double)v[k];
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
      sm[k] += __shfl_xor_sync(0xffffffffu, sm[k], d);
      sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], d);
    }
  }
  // one set of atomics per block (per-warp atomics on 13 shared addresses cost 58 us at 200 k points)
  __shared__ double s_sm[8][3], s_sq[8][3];
  __shared__ float s_lo[8][3], s_hi[8][3];
  __shared__ unsigned s_cnt[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_cnt[warp] = cnt;
    for (int k = 0; k < 3; ++k) { s_sm[warp][k] = sm[k]; s_sq[warp][k] = sq[k]; s_lo[warp][k] = lo[k]; s_hi[warp][k] = hi[k]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      cnt += s_cnt[w];
      for (int k = 0; k < 3; ++k) {
        sm[k] += s_sm[w][k]; sq[k] += s_sq[w][k];
        lo[k] = fminf(lo[k], s_lo[w][k]); hi[k] = fmaxf(hi[k], s_hi[w][k]);
      }
    }
    if (cnt) {
      atomicAdd(&st->count, (unsigned long long)cnt);
      for (int k = 0; k < 3; ++k) {
        atomicMin(&st->lo[k], f2ord(lo[k])); atomicMax(&st->hi[k], f2ord(hi[k]));
        atomicAdd(&st->sum[k], sm[k]); atomicAdd(&st->sumsq[k], sq[k]);
      }
    }
  }
}

__device__ __forceinline__ int ceil_log2_i(long long v) {
  int b = 0;
  while ((1ll << b) < v) ++b;
  return b;
}

// grid = robust box (mean +- 4 sigma clipped to the bounding box), cell edge doubled until the Morton key
// space 2^(2*bxy + bz) fits the bitmap
__global__ void k_nn_params(const NNStats* __restrict__ st, float cell, int max_bits, NNGrid* __restrict__ g) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
  float amax = 0.f;
  if (st->count) {
    const double n = (double)st->count;
    for (int k = 0; k < 3; ++k) {
      const float bl = ord2f(st->lo[k]), bh = ord2f(st->hi[k]);
      const double mean = st->sum[k] / n;
      double var = st->sumsq[k] / n - mean * mean;
      const double sd = var > 0.0 ? sqrt(var) : 0.0;
      lo[k] = fmaxf(bl, (float)(mean - 4.0 * sd));
      hi[k] = fminf(bh, (float)(mean + 4.0 * sd));
      if (!(hi[k] >= lo[k])) { lo[k] = bl; hi[k] = bl; }
      amax = fmaxf(amax, fmaxf(fabsf(bl), fabsf(bh)));
    }
  }
  float h = cell;
  int bxy = 0, bz = 0;
  for (int it = 0; it < 64; ++it) {
    const long long cx = (long long)fminf(floorf((hi[0] - lo[0]) / h) + 1.f, 1.0e9f);
    const long long cy = (long long)fminf(floorf((hi[1] - lo[1]) / h) + 1.f, 1.0e9f);
    const long long cz = (long long)fminf(floorf((hi[2] - lo[2]) / h) + 1.f, 1.0e9f);
    bxy = ceil_log2_i(cx > cy ? cx : cy);
    bz = ceil_log2_i(cz);
    if (bz > bxy) bxy = bz;
    if (2 * bxy + bz <= max_bits && bxy <= 13 && bz <= 10) break;
    h *= 2.f;
  }
  g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
  g->h = h;
  g->eps = 1e-3f * h + 8.f * 1.2e-7f * amax;
  g->bxy = bxy; g->bz = bz;
  g->n_words = (int)(((1ll << (2 * bxy + bz)) >> 6) + 1);
}

// ---- Morton key with unequal bit counts: the low bz levels interleave (x,y,z), the upper bxy-bz levels (x,y)
__device__ __forceinline__ unsigned part1by2(unsigned x) {   // 10 bits -> every third bit
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
__device__ __forceinline__ unsigned part1by1(unsigned x) {   // 16 bits -> every second bit
  x &= 0xffffu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
__device__ __forceinline__ unsigned compact1by2(unsigned x) {
  x &= 0x09249249u;
  x = (x ^ (x >> 2)) & 0x030C30C3u;
  x = (x ^ (x >> 4)) & 0x0300F00Fu;
  x = (x ^ (x >> 8)) & 0x030000FFu;
  x = (x ^ (x >> 16)) & 0x3ffu;
  return x;
}
__device__ __forceinline__ unsigned compact1by1(unsigned x) {
  x &= 0x55555555u;
  x = (x ^ (x >> 1)) & 0x33333333u;
  x = (x ^ (x >> 2)) & 0x0F0F0F0Fu;
  x = (x ^ (x >> 4)) & 0x00FF00FFu;
  x = (x ^ (x >> 8)) & 0xffffu;
  return x;
}
// dilated x component (y: shift the result left by 1; z: part1by2(cz) << 2)
__device__ __forceinline__ unsigned dil_xy(unsigned c, int bz) {
  const unsigned lowmask = (1u << bz) - 1u;
  return part1by2(c & lowmask) | (part1by1(c >> bz) << (3 * bz));
}
__device__ __forceinline__ unsigned nn_key(const NNGrid& g, int cx, int cy, int cz) {
  return dil_xy((unsigned)cx, g.bz) | (dil_xy((unsigned)cy, g.bz) << 1) | (part1by2((unsigned)cz) << 2);
}
__device__ __forceinline__ void nn_unkey(const NNGrid& g, unsigned key, int& cx, int& cy, int& cz) {
  const unsigned low = key & ((1u << (3 * g.bz)) - 1u), high = key >> (3 * g.bz);
  cx = (int)(compact1by2(low) | (compact1by1(high) << g.bz));
  cy = (int)(compact1by2(low >> 1) | (compact1by1(high >> 1) << g.bz));
  cz = (int)compact1by2(low >> 2);
}

// cell of a point; points outside the grid (or non-finite) land in the nearest boundary cell
__device__ __forceinline__ void nn_cell(const NNGrid& g, float x, float y, float z, int& cx, int& cy, int& cz) {
  const int nxy = (1 << g.bxy) - 1, nz = (1 << g.bz) - 1;
  cx = min(max(__float2int_rd(__fdiv_rn(x - g.ox, g.h)), 0), nxy);
  cy = min(max(__float2int_rd(__fdiv_rn(y - g.oy, g.h)), 0), nxy);
  cz = min(max(__float2int_rd(__fdiv_rn(z - g.oz, g.h)), 0), nz);
}

struct NNCloud {
  const float* pts;              // [n,3]
  int n;
  unsigned* keys;                // [n] Morton key, then occupied-cell rank
  unsigned long long* bitmap;    // [kNNMaxWords]
  int* word_prefix;              // [kNNMaxWords]
  int* count;                    // [n+1] zeroed
  int* slot;                     // [n]
  int* cell_start;               // [n+2]
  float4* sorted;                // [n] xyz + original index bits
  int* n_cells_occ;              // [1]
};

__device__ __forceinline__ int nn_rank(const unsigned long long* __restrict__ bitmap, const int* __restrict__ prefix,
                                       unsigned key) {
  const unsigned w = key >> 6;
  const unsigned long long below = (1ull << (key & 63u)) - 1ull;
  return __ldg(prefix + w) + __popcll(__ldg(bitmap + w) & below);
}

__global__ void __launch_bounds__(256)
k_nn_clear(NNCloud c0, NNCloud c1, const NNGrid* __restrict__ gp) {
  const int nw = gp->n_words;
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += gridDim.x * blockDim.x) c.bitmap[i] = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= c.n; i += gridDim.x * blockDim.x) c.count[i] = 0;
}

__global__ void __launch_bounds__(256)
k_nn_mark(NNCloud c0, NNCloud c1, const NNGrid* __restrict__ gp) {
  const NNGrid g = *gp;
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    int cx, cy, cz;
    nn_cell(g, __ldg(c.pts + 3 * (size_t)i), __ldg(c.pts + 3 * (size_t)i + 1),
            __ldg(c.pts + 3 * (size_t)i + 2), cx, cy, cz);
    const unsigned key = nn_key(g, cx, cy, cz);
    c.keys[i] = key;
    atomicOr(c.bitmap + (key >> 6), 1ull << (key & 63u));
  }
}

struct MapPopc64 {
  const unsigned long long* p;
  __device__ int operator()(int i) const { return __popcll(p[i]); }
};

__global__ void __launch_bounds__(256)
k_nn_rank(NNCloud c0, NNCloud c1) {
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    const int rank = nn_rank(c.bitmap, c.word_prefix, c.keys[i]);
    c.keys[i] = (unsigned)rank;  // keys now hold the occupied-cell rank
    c.slot[i] = atomicAdd(c.count + rank, 1);
  }
}

__global__ void __launch_bounds__(256)
k_nn_fill(NNCloud c0, NNCloud c1) {
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    const int pos = c.cell_start[c.keys[i]] + c.slot[i];
    c.sorted[pos] = make_float4(__ldg(c.pts + 3 * (size_t)i), __ldg(c.pts + 3 * (size_t)i + 1),
                                __ldg(c.pts + 3 * (size_t)i + 2), __int_as_float(i));
  }
}

// ---- search ---------------------------------------------------------------------------------------------
struct NNQuery {
  float x, y, z;        // the query
  float px, py, pz;     // the query projected onto the grid box (all box bounds use this)
  float best;
  int best_i;
};

__device__ __forceinline__ void nn_scan(const float4* __restrict__ rs, int s, int e, NNQuery& q) {
  for (int j = s; j < e; ++j) {
    const float4 c = __ldg(rs + j);
    const float dx = c.x - q.x, dy = c.y - q.y, dz = c.z - q.z;
    const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const int ci = __float_as_int(c.w);
    if (d < q.best || (d == q.best && ci < q.best_i)) { q.best = d; q.best_i = ci; }
  }
}

// squared distance from the projected query to the level-`lvl` node with node coordinates (bx,by,bz_), shrunk by eps
__device__ __forceinline__ float nn_box_d2(const NNGrid& g, const NNQuery& q, int lvl, int bx, int by, int bz_) {
  const float H = g.h * (float)(1 << lvl);
  const float lx = g.ox + (float)bx * H, ly = g.oy + (float)by * H, lz = g.oz + (float)bz_ * H;
  const float dx = fmaxf(fmaxf(lx - q.px, q.px - (lx + H)) - g.eps, 0.f);
  const float dy = fmaxf(fmaxf(ly - q.py, q.py - (ly + H)) - g.eps, 0.f);
  const float dz = fmaxf(fmaxf(lz - q.pz, q.pz - (lz + H)) - g.eps, 0.f);
  return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ int nn_level_bits(const NNGrid& g, int lvl) {
  return 2 * min(lvl, g.bxy) + min(lvl, g.bz);
}

struct NNRef {
  const unsigned long long* __restrict__ bitmap;
  const int* __restrict__ prefix;
  const int* __restrict__ cstart;
  const float4* __restrict__ rs;
};

// point range of the node whose first key is `base` (level lvl)
__device__ __forceinline__ void nn_node_range(const NNGrid& g, const NNRef& r, unsigned base, int lvl, int& s, int& e) {
  const unsigned size = 1u << nn_level_bits(g, lvl);
  s = __ldg(r.cstart + nn_rank(r.bitmap, r.prefix, base));
  e = __ldg(r.cstart + nn_rank(r.bitmap, r.prefix, base + size));
}

// key bit that selects the upper half of a level-(cl+1) node along x (y: the next bit) and along z
__device__ __forceinline__ unsigned nn_xbit(const NNGrid& g, int cl) {
  return 1u << (cl < g.bz ? 3 * cl : 3 * g.bz + 2 * (cl - g.bz));
}

// depth-first descent of the nodes on `stack`, nearest child first, pruned by the current best.
// Inlined at its single call site so that the query state stays in registers.
__device__ __forceinline__ void nn_descend(const NNGrid& g, const NNRef& r, NNQuery& q, int qcx, int qcy, int qcz,
                                           unsigned* stack, int sp) {
  while (sp > 0) {
    const unsigned ent = stack[--sp];
    const unsigned base = ent & ((1u << 28) - 1u);      // keys have up to 28 bits, levels <= 13
    const int lvl = (int)(ent >> 28);
    int cx, cy, cz;
    nn_unkey(g, base, cx, cy, cz);
    if (nn_box_d2(g, q, lvl, cx >> lvl, cy >> lvl, cz >> lvl) > q.best) continue;
    int s, e;
    nn_node_range(g, r, base, lvl, s, e);
    if (s == e) continue;
    if (e - s <= kNNLeaf || lvl == 0) { nn_scan(r.rs, s, e, q); continue; }
    // children at level lvl-1: an axis splits when it still has a bit at that level
    const int cl = lvl - 1;
    const bool sxy = cl < g.bxy, sz = cl < g.bz;
    const unsigned bx = nn_xbit(g, cl), bzb = 4u << (3 * cl);
    const float Hc = g.h * (float)(1 << cl);
    const float lx = g.ox + (float)cx * g.h, ly = g.oy + (float)cy * g.h, lz = g.oz + (float)cz * g.h;   // node origin
    // per-axis squared distance to the lower / upper half
    float ax[2], ay[2], az[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float ox_ = lx + (float)hf * Hc, oy_ = ly + (float)hf * Hc, oz_ = lz + (float)hf * Hc;
      const float tx = fmaxf(fmaxf(ox_ - q.px, q.px - (ox_ + Hc)) - g.eps, 0.f);
      const float ty = fmaxf(fmaxf(oy_ - q.py, q.py - (oy_ + Hc)) - g.eps, 0.f);
      const float tz = fmaxf(fmaxf(oz_ - q.pz, q.pz - (oz_ + Hc)) - g.eps, 0.f);
      ax[hf] = tx * tx; ay[hf] = ty * ty; az[hf] = tz * tz;
    }
    if (!sxy) { ax[1] = ax[0]; ay[1] = ay[0]; }      // axis does not split: the child spans what the parent spans
    if (!sz) {                                       // (its box is the parent's: use the parent's extent)
      const float tz = fmaxf(fmaxf(lz - q.pz, q.pz - (lz + 2.f * Hc)) - g.eps, 0.f);
      az[0] = az[1] = tz * tz;
    }
    const int nx_ = sxy && ax[1] < ax[0], ny_ = sxy && ay[1] < ay[0], nz_ = sz && az[1] < az[0];   // nearer half
    const int nchild = (sxy ? 4 : 1) * (sz ? 2 : 1);
    // pushed farthest-first so that the nearest child is popped first
    for (int m = nchild - 1; m >= 0; --m) {
      int ix = 0, iy = 0, iz = 0, mm = m;
      if (sxy) { ix = (mm & 1) ^ nx_; iy = ((mm >> 1) & 1) ^ ny_; mm >>= 2; }
      if (sz) iz = (mm & 1) ^ nz_;
      if ((ix ? ax[1] : ax[0]) + (iy ? ay[1] : ay[0]) + (iz ? az[1] : az[0]) > q.best) continue;
      const unsigned ckey = base | (ix ? bx : 0u) | (iy ? bx << 1 : 0u) | (iz ? bzb : 0u);
      if (sp < 96) stack[sp++] = ckey | ((unsigned)cl << 28);
      else {   // cannot happen (depth <= 13, <= 7 pending siblings per level); scan rather than drop
        int cs, ce;
        nn_node_range(g, r, ckey, cl, cs, ce);
        nn_scan(r.rs, cs, ce, q);
      }
    }
  }
}

// visiting order of the 3x3x3 nodes: centre, the 6 face neighbours, the 12 edge neighbours, the 8 corners
__constant__ unsigned char kNNOrder[27] = {13, 4, 10, 12, 14, 16, 22, 1, 3, 5, 7, 9, 11, 15, 17, 19, 21, 23, 25,
                                           0, 2, 6, 8, 18, 20, 24, 26};

// One thread per query, queries taken in Morton order so that the lanes of a warp walk the same nodes.
// Both directions run in one launch (blockIdx.y).  radius2: +inf, or the squared search radius.
__global__ void __launch_bounds__(128)
k_nn_search(NNCloud c0, NNCloud c1, const NNGrid* __restrict__ gp, float* __restrict__ dist0,
            int32_t* __restrict__ idx0, float* __restrict__ dist1, int32_t* __restrict__ idx1, float radius2) {
  const NNGrid g = *gp;
  const NNCloud& qc = blockIdx.y == 0 ? c0 : c1;
  const NNCloud& rc = blockIdx.y == 0 ? c1 : c0;
  float* __restrict__ dist = blockIdx.y == 0 ? dist0 : dist1;
  int32_t* __restrict__ idx = blockIdx.y == 0 ? idx0 : idx1;
  NNRef r;
  r.bitmap = rc.bitmap; r.prefix = rc.word_prefix; r.cstart = rc.cell_start; r.rs = rc.sorted;
  const int nxy = 1 << g.bxy, nz = 1 << g.bz;
  const int top = g.bxy;                                   // level at which one node covers the grid
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < qc.n; t += gridDim.x * blockDim.x) {
    const float4 p = qc.sorted[t];
    const int qi = __float_as_int(p.w);
    NNQuery q;
    q.x = p.x; q.y = p.y; q.z = p.z;
    q.best = radius2; q.best_i = -1;
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {   // the reference's comparisons all fail on NaN
      dist[qi] = 1e20f; idx[qi] = -1;
      continue;
    }
    const float ex = g.h * (float)nxy, ez = g.h * (float)nz;
    q.px = fminf(fmaxf(p.x, g.ox), g.ox + ex);
    q.py = fminf(fmaxf(p.y, g.oy), g.oy + ex);
    q.pz = fminf(fmaxf(p.z, g.oz), g.oz + ez);
    int cx, cy, cz;
    nn_cell(g, p.x, p.y, p.z, cx, cy, cz);
    unsigned stack[96];
    // start at the finest level whose node around the query holds a reference point: in sparse regions the
    // 27 empty probes of every finer level are skipped (any start level is exact, see the slab bound below)
    int lvl0 = 0;
    {
      const unsigned kq = nn_key(g, cx, cy, cz);
      for (; lvl0 < top; ++lvl0) {
        const unsigned size = 1u << nn_level_bits(g, lvl0);
        int s, e;
        nn_node_range(g, r, kq & ~(size - 1u), lvl0, s, e);
        if (e > s) break;
      }
    }
    for (int lvl = lvl0; lvl <= top; ++lvl) {
      const float H = g.h * (float)(1 << lvl);
      const int kx = cx >> lvl, ky = cy >> lvl, kz = lvl < g.bz ? cz >> lvl : 0;
      const int mxy = max(nxy >> lvl, 1), mz = max(nz >> lvl, 1);
      // per-axis tables for the three offsets: dilated key component, squared box distance, validity
      unsigned KX[3], KY[3], KZ[3];
      float DX[3], DY[3], DZ[3];
      bool VX[3], VY[3], VZ[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int x = kx + d - 1, y = ky + d - 1, z = kz + d - 1;
        VX[d] = x >= 0 && x < mxy; VY[d] = y >= 0 && y < mxy; VZ[d] = z >= 0 && z < mz;
        KX[d] = dil_xy((unsigned)(max(x, 0) << lvl), g.bz);
        KY[d] = dil_xy((unsigned)(max(y, 0) << lvl), g.bz) << 1;
        KZ[d] = part1by2((unsigned)(max(z, 0) << lvl)) << 2;
        const float lx = g.ox + (float)x * H, ly = g.oy + (float)y * H, lz = g.oz + (float)z * H;
        const float tx = fmaxf(fmaxf(lx - q.px, q.px - (lx + H)) - g.eps, 0.f);
        const float ty = fmaxf(fmaxf(ly - q.py, q.py - (ly + H)) - g.eps, 0.f);
        const float tz = fmaxf(fmaxf(lz - q.pz, q.pz - (lz + H)) - g.eps, 0.f);
        DX[d] = tx * tx; DY[d] = ty * ty; DZ[d] = tz * tz;
      }
      // first level: the query's own node first (its points prune most neighbours); later levels: the centre
      // node is the parent of the previous level's node and lies inside the slab already examined -- skip it.
      // The 27-node loop stays rolled (register tables read through selects) to keep the code small: the fully
      // unrolled version ran 60 % slower on instruction fetch.
#pragma unroll 1
      for (int t = lvl == lvl0 ? 0 : 1; t < 27; ++t) {
        const int i = kNNOrder[t];
        const int iz = i / 9, iy = (i - 9 * iz) / 3, ix = i - 9 * iz - 3 * iy;
        const bool v = (ix == 0 ? VX[0] : ix == 1 ? VX[1] : VX[2]) && (iy == 0 ? VY[0] : iy == 1 ? VY[1] : VY[2]) &&
                       (iz == 0 ? VZ[0] : iz == 1 ? VZ[1] : VZ[2]);
        if (!v) continue;
        const float d2 = (ix == 0 ? DX[0] : ix == 1 ? DX[1] : DX[2]) + (iy == 0 ? DY[0] : iy == 1 ? DY[1] : DY[2]) +
                         (iz == 0 ? DZ[0] : iz == 1 ? DZ[1] : DZ[2]);
        if (d2 > q.best) continue;                               // cannot hold a better (or equal, lower-index) point
        const unsigned base = (ix == 0 ? KX[0] : ix == 1 ? KX[1] : KX[2]) | (iy == 0 ? KY[0] : iy == 1 ? KY[1] : KY[2]) |
                              (iz == 0 ? KZ[0] : iz == 1 ? KZ[1] : KZ[2]);
        int s, e;
        if (lvl == 0) {
          const unsigned long long w = __ldg(r.bitmap + (base >> 6));
          if (!((w >> (base & 63u)) & 1ull)) continue;
          const int rk = __ldg(r.prefix + (base >> 6)) + __popcll(w & ((1ull << (base & 63u)) - 1ull));
          s = __ldg(r.cstart + rk); e = __ldg(r.cstart + rk + 1);
        } else {
          nn_node_range(g, r, base, lvl, s, e);
          if (s == e) continue;
          if (e - s > kNNLeaf) {                              // big node: depth-first, nearest child first
            stack[0] = base | ((unsigned)lvl << 28);
            nn_descend(g, r, q, cx, cy, cz, stack, 1);
            continue;
          }
        }
        nn_scan(r.rs, s, e, q);
      }
      // everything stored outside the 3x3x3 slab is at least `bound` away (sides cut by the grid edge are open:
      // nothing is stored beyond them)
      float bound = INFINITY;
      if (kx - 1 >= 0) bound = fminf(bound, q.px - (g.ox + (float)(kx - 1) * H));
      if (kx + 1 < mxy) bound = fminf(bound, (g.ox + (float)(kx + 2) * H) - q.px);
      if (ky - 1 >= 0) bound = fminf(bound, q.py - (g.oy + (float)(ky - 1) * H));
      if (ky + 1 < mxy) bound = fminf(bound, (g.oy + (float)(ky + 2) * H) - q.py);
      if (kz - 1 >= 0) bound = fminf(bound, q.pz - (g.oz + (float)(kz - 1) * H));
      if (kz + 1 < mz) bound = fminf(bound, (g.oz + (float)(kz + 2) * H) - q.pz);
      bound -= g.eps;
      if (bound == INFINITY) break;                               // whole grid examined
      if (bound > 0.f && q.best < bound * bound) break;
    }
    if (q.best_i < 0) q.best = 1e20f;                             // radius-limited search found nothing
    dist[qi] = q.best;
    idx[qi] = q.best_i;
  }
}

// ---- warp-cooperative search ---------------------------------------------------------------------------------
// One WARP per query (queries in Morton order, so neighbouring warps walk the same nodes and share them in L1/L2).
// Control flow is uniform across the warp; the lanes share the work of a query instead of each owning one:
//   * start level: lane l looks up the population of the level-l node around the query (two rank lookups per lane,
//     all levels at once); the search starts at the finest level whose node holds >= kNNWStart points, so that a leaf
//     scan fills the warp (0.5 m lidar cells hold ~4 points: scanning them one by one leaves 28 lanes idle);
//   * 3x3x3 probe: lane t < 27 owns neighbour node t -- box distance, key, point range -- and the warp then visits the
//     surviving nodes nearest first (redux.sync min over the box distances + ballot), re-pruning after every scan;
//   * leaf scan: lanes read CONSECUTIVE float4 of the sorted reference cloud (one 512-byte coalesced request per step),
//     each lane keeps its own (best, index); the warp-wide bound is one redux.sync min over the distance bits;
//   * crowded nodes (> kNNWLeaf points) are descended depth-first with a per-warp stack in shared memory; the <= 8
//     children are evaluated by 8 lanes and pushed farthest first;
//   * the result is the 64-bit min over the lanes of (distance bits << 32 | index): ties resolve to the lowest index,
//     the reference's rule (chamfer3D.cu:61-69 scans j ascending with a strict <).
// Exactness is the per-thread kernel's: the same box bounds (projection onto the grid box, eps slack) and the same slab
// bound end the level loop; only the order in which candidates are tested differs, and (best, index) is order-free.
constexpr int kNNWLeaf = 256;    // nodes up to this many points are scanned (8 coalesced steps), larger ones descended
constexpr int kNNWStart = 24;    // start at the finest level whose centre node holds at least this many points
constexpr int kNNWWarps = 4;

__device__ __forceinline__ void nnw_scan(const float4* __restrict__ rs, int s, int e, int lane, const NNQuery& qq,
                                         float& best, int& best_i) {
  for (int j = s + lane; j < e; j += 32) {
    const float4 c = __ldg(rs + j);
    const float dx = c.x - qq.x, dy = c.y - qq.y, dz = c.z - qq.z;
    const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
    const int ci = __float_as_int(c.w);
    if (d < best || (d == best && ci < best_i)) { best = d; best_i = ci; }
  }
}

__global__ void __launch_bounds__(32 * kNNWWarps)
k_nn_search_warp(NNCloud c0, NNCloud c1, const NNGrid* __restrict__ gp, float* __restrict__ dist0,
                 int32_t* __restrict__ idx0, float* __restrict__ dist1, int32_t* __restrict__ idx1, float radius2) {
  __shared__ unsigned s_stack[kNNWWarps][96];
  const NNGrid g = *gp;
  const NNCloud& qc = blockIdx.y == 0 ? c0 : c1;
  const NNCloud& rc = blockIdx.y == 0 ? c1 : c0;
  float* __restrict__ dist = blockIdx.y == 0 ? dist0 : dist1;
  int32_t* __restrict__ idx = blockIdx.y == 0 ? idx0 : idx1;
  NNRef r;
  r.bitmap = rc.bitmap; r.prefix = rc.word_prefix; r.cstart = rc.cell_start; r.rs = rc.sorted;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned* stack = s_stack[wib];
  const unsigned FULL = 0xffffffffu;
  const int nxy = 1 << g.bxy, nz = 1 << g.bz;
  const int top = g.bxy;
  const float ex = g.h * (float)nxy, ez = g.h * (float)nz;
  for (int t = blockIdx.x * kNNWWarps + wib; t < qc.n; t += gridDim.x * kNNWWarps) {
    const float4 p = qc.sorted[t];                       // same address in every lane: one broadcast load
    const int qi = __float_as_int(p.w);
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) {   // the reference's comparisons all fail on NaN
      if (lane == 0) { dist[qi] = 1e20f; idx[qi] = -1; }
      continue;
    }
    NNQuery q;
    q.x = p.x; q.y = p.y; q.z = p.z;
    q.px = fminf(fmaxf(p.x, g.ox), g.ox + ex);
    q.py = fminf(fmaxf(p.y, g.oy), g.oy + ex);
    q.pz = fminf(fmaxf(p.z, g.oz), g.oz + ez);
    float best = radius2;            // lane-private running minimum
    int best_i = -1;
    float best_w = radius2;          // warp-wide bound used for pruning (uniform)
    int cx, cy, cz;
    nn_cell(g, p.x, p.y, p.z, cx, cy, cz);
    const unsigned kq = nn_key(g, cx, cy, cz);
    // ---- start level: lane l counts the level-l node around the query
    int lvl0;
    {
      int cnt = 0;
      if (lane <= top) {
        const unsigned size = 1u << nn_level_bits(g, lane);
        int s, e;
        nn_node_range(g, r, kq & ~(size - 1u), lane, s, e);
        cnt = e - s;
      }
      const unsigned m = __ballot_sync(FULL, lane <= top && cnt >= kNNWStart);
      lvl0 = m ? __ffs(m) - 1 : top;
    }
    for (int lvl = lvl0; lvl <= top; ++lvl) {
      const float H = g.h * (float)(1 << lvl);
      const int kx = cx >> lvl, ky = cy >> lvl, kz = lvl < g.bz ? cz >> lvl : 0;
      const int mxy = max(nxy >> lvl, 1), mz = max(nz >> lvl, 1);
      // ---- probe: lane t27 owns one of the 3x3x3 nodes
      int ns = 0, ne = 0;
      unsigned nbase = 0u, nkey = 0xffffffffu;      // nkey: box-distance bits while the node is still to be visited
      if (lane < 27 && !(lvl != lvl0 && lane == 13)) {   // later levels: the centre is the parent of the slab already examined
        const int iz = lane / 9, iy = (lane - 9 * iz) / 3, ix = lane - 9 * iz - 3 * iy;
        const int x = kx + ix - 1, y = ky + iy - 1, z = kz + iz - 1;
        if (x >= 0 && x < mxy && y >= 0 && y < mxy && z >= 0 && z < mz) {
          const float d2 = nn_box_d2(g, q, lvl, x, y, z);
          if (d2 <= best_w) {
            nbase = nn_key(g, x << lvl, y << lvl, z << lvl);
            nn_node_range(g, r, nbase, lvl, ns, ne);
            if (ne > ns) nkey = __float_as_uint(d2);
          }
        }
      }
      // ---- visit the surviving nodes nearest first
      while (true) {
        if (nkey != 0xffffffffu && __uint_as_float(nkey) > best_w) nkey = 0xffffffffu;
        const unsigned mn = __reduce_min_sync(FULL, nkey);
        if (mn == 0xffffffffu) break;
        const int src = __ffs(__ballot_sync(FULL, nkey == mn)) - 1;
        const int s = __shfl_sync(FULL, ns, src), e = __shfl_sync(FULL, ne, src);
        const unsigned base = __shfl_sync(FULL, nbase, src);
        if (lane == src) nkey = 0xffffffffu;
        if (e - s <= kNNWLeaf || lvl == 0) {
          nnw_scan(r.rs, s, e, lane, q, best, best_i);
        } else {
          // ---- depth-first descent, nearest child first (uniform control flow; 8 lanes evaluate the children)
          int sp = 0;
          if (lane == 0) stack[0] = base | ((unsigned)lvl << 28);
          sp = 1;
          __syncwarp();
          while (sp > 0) {
            const unsigned ent = stack[--sp];
            __syncwarp();
            const unsigned nb = ent & ((1u << 28) - 1u);
            const int nl = (int)(ent >> 28);
            int bx_, by_, bz_;
            nn_unkey(g, nb, bx_, by_, bz_);
            if (nn_box_d2(g, q, nl, bx_ >> nl, by_ >> nl, bz_ >> nl) > best_w) continue;
            int cs, ce;
            nn_node_range(g, r, nb, nl, cs, ce);
            if (cs == ce) continue;
            if (ce - cs <= kNNWLeaf || nl == 0) {
              nnw_scan(r.rs, cs, ce, lane, q, best, best_i);
              best_w = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(best)));
              continue;
            }
            const int cl = nl - 1;
            const bool sxy = cl < g.bxy, sz = cl < g.bz;
            const unsigned xb = nn_xbit(g, cl), zb = 4u << (3 * cl);
            const float Hc = g.h * (float)(1 << cl);
            const float lx = g.ox + (float)bx_ * g.h, ly = g.oy + (float)by_ * g.h, lz = g.oz + (float)bz_ * g.h;
            const int nchild = (sxy ? 4 : 1) * (sz ? 2 : 1);
            unsigned ckey = 0u, ckd = 0xffffffffu;
            if (lane < nchild) {
              int ix = 0, iy = 0, iz = 0, mm = lane;
              if (sxy) { ix = mm & 1; iy = (mm >> 1) & 1; mm >>= 2; }
              if (sz) iz = mm & 1;
              const float ox_ = lx + (float)ix * Hc, oy_ = ly + (float)iy * Hc, oz_ = lz + (float)iz * Hc;
              const float wxy = sxy ? Hc : 2.f * Hc, wz = sz ? Hc : 2.f * Hc;      // an axis that does not split keeps the parent's extent
              const float tx = fmaxf(fmaxf(ox_ - q.px, q.px - (ox_ + wxy)) - g.eps, 0.f);
              const float ty = fmaxf(fmaxf(oy_ - q.py, q.py - (oy_ + wxy)) - g.eps, 0.f);
              const float tz = fmaxf(fmaxf(oz_ - q.pz, q.pz - (oz_ + wz)) - g.eps, 0.f);
              const float cd = tx * tx + ty * ty + tz * tz;
              if (cd <= best_w) {
                ckey = nb | (ix ? xb : 0u) | (iy ? xb << 1 : 0u) | (iz ? zb : 0u);
                int s2, e2;
                nn_node_range(g, r, ckey, cl, s2, e2);
                if (e2 > s2) ckd = __float_as_uint(cd);
              }
            }
            // rank of this child among the valid ones by (distance, lane): nearest = rank 0
            const unsigned vmask = __ballot_sync(FULL, ckd != 0xffffffffu);
            const int nvalid = __popc(vmask);
            int rank = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const unsigned kj = __shfl_sync(FULL, ckd, j);
              rank += (kj < ckd || (kj == ckd && j < lane)) ? 1 : 0;
            }
            if (ckd != 0xffffffffu && sp + nvalid <= 96) stack[sp + (nvalid - 1 - rank)] = ckey | ((unsigned)cl << 28);
            if (sp + nvalid <= 96) sp += nvalid;
            else {   // cannot happen (depth <= 13, <= 7 pending siblings per level); scan rather than drop
              nnw_scan(r.rs, cs, ce, lane, q, best, best_i);
              best_w = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(best)));
            }
            __syncwarp();
          }
        }
        best_w = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(best)));
      }
      // everything stored outside the 3x3x3 slab is at least `bound` away (sides cut by the grid edge are open)
      float bound = INFINITY;
      if (kx - 1 >= 0) bound = fminf(bound, q.px - (g.ox + (float)(kx - 1) * H));
      if (kx + 1 < mxy) bound = fminf(bound, (g.ox + (float)(kx + 2) * H) - q.px);
      if (ky - 1 >= 0) bound = fminf(bound, q.py - (g.oy + (float)(ky - 1) * H));
      if (ky + 1 < mxy) bound = fminf(bound, (g.oy + (float)(ky + 2) * H) - q.py);
      if (kz - 1 >= 0) bound = fminf(bound, q.pz - (g.oz + (float)(kz - 1) * H));
      if (kz + 1 < mz) bound = fminf(bound, (g.oz + (float)(kz + 2) * H) - q.pz);
      bound -= g.eps;
      if (bound == INFINITY) break;
      if (bound > 0.f && best_w < bound * bound) break;
    }
    // ---- 64-bit min over the lanes: (distance bits, index); an untouched lane holds (radius2, -1) and sorts last
    unsigned long long pk = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)best_i;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(FULL, pk, d);
      pk = o < pk ? o : pk;
    }
    if (lane == 0) {
      const int bi = (int)(unsigned)(pk & 0xffffffffull);
      dist[qi] = bi < 0 ? 1e20f : __uint_as_float((unsigned)(pk >> 32));
      idx[qi] = bi;
    }
  }
}

__global__ void k_nn_fill_empty(float* dist, int32_t* idx, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    dist[i] = 1e20f;
    idx[i] = -1;
  }
}

// grad of sum_i g[i]*dist[i] wrt both clouds, one direction per launch
__global__ void __launch_bounds__(256)
k_chamfer_grad(const float* __restrict__ a, int na, const float* __restrict__ b,
               const int32_t* __restrict__ idx, const float* __restrict__ gd, float* ga, float* gb) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) {
    const int j = idx[i];
    if (j < 0) continue;
    const float g = gd[i] * 2.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float t = g * (a[3 * (size_t)i + k] - b[3 * (size_t)j + k]);
      atomicAdd(ga + 3 * (size_t)i + k, t);
      atomicAdd(gb + 3 * (size_t)j + k, -t);
    }
  }
}

static size_t nn_cloud_bytes(int n, long long kNNMaxWords) {
  size_t m = (size_t)(n > 0 ? n : 1);
  size_t b = 0;
  b += align_up(m * sizeof(unsigned), 256);                       // keys
  b += align_up((size_t)kNNMaxWords * sizeof(unsigned long long), 256);   // bitmap
  b += align_up((size_t)kNNMaxWords * sizeof(int), 256);          // word prefix
  b += align_up((m + 1) * sizeof(int), 256);                      // count
  b += align_up(m * sizeof(int), 256);                            // slot
  b += align_up((m + 2) * sizeof(int), 256);                      // cell_start
  b += align_up(m * sizeof(float4), 256);                         // sorted
  b += 256;                                                       // n_cells_occ
  b += ScanScratch::bytes(kNNMaxWords) + ScanScratch::bytes((long long)m + 2);
  return b;
}

}  // namespace himo

using namespace himo;

extern "C" size_t himo_chamfer_workspace_bytes(int n0, int n1) {
  if (n0 < 0 || n1 < 0) return 0;
  const long long mw = nn_max_words((long long)n0 + n1);
  return nn_cloud_bytes(n0, mw) + nn_cloud_bytes(n1, mw) + 4096 + 16 * 256;
}

// The warp-per-query kernel returns identical results but is SLOWER on every measured cloud (B200, round 2,
// profiles/r02_nn_warp_vs_thread.jsonl: lidar 100 k 0.78 vs 0.65 ms, uniform 100 k 0.31 vs 0.18 ms, lidar 1 M 13.7 vs
// 3.85 ms): the search is bound by chains of dependent loads (bitmap word -> prefix -> cell start -> points), and a
// warp that owns one query has one such chain in flight where a warp of 32 single-thread queries has 32.  Coalesced
// scans and shuffle reductions do not pay for a 32x loss of memory-level parallelism.  Kept as an A/B knob.
static int g_nn_warp = 0;
// A/B knob: 1 selects the warp-cooperative search kernel (one warp per query) instead of one thread per query.
extern "C" int himo_chamfer_set_warp_search(int enable) { g_nn_warp = enable ? 1 : 0; return HIMO_OK; }

static int nn_forward_impl(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                           float* dist1, int32_t* idx0, int32_t* idx1, float cell_size, float radius2,
                           void* workspace, size_t workspace_bytes, void* stream_) {
  if (n0 < 0 || n1 < 0) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if ((n0 > 0 && (!pc0 || !dist0 || !idx0)) || (n1 > 0 && (!pc1 || !dist1 || !idx1))) return HIMO_ERR_ARG;
  if (n0 == 0 || n1 == 0) {
    // empty other cloud: the reference's scan never updates (best=1e20, best_i=-1)
    if (n0 > 0) { k_nn_fill_empty<<<ceil_div(n0, 256), 256, 0, stream>>>(dist0, idx0, n0); HIMO_LAUNCH_RET(); }
    if (n1 > 0) { k_nn_fill_empty<<<ceil_div(n1, 256), 256, 0, stream>>>(dist1, idx1, n1); HIMO_LAUNCH_RET(); }
    return HIMO_OK;
  }
  // default finest cell: about the point spacing -- 0.5 m up to a few hundred thousand points, 0.125 m beyond
  // (measured: 0.5 m is 4-20 % faster at 100 k points per cloud; coarser levels come for free from the key)
  if (!(cell_size > 0.f)) cell_size = ((long long)n0 + n1 < 400000) ? 0.5f : 0.125f;
  const long long kNNMaxWords = nn_max_words((long long)n0 + n1);
  Arena A(workspace, workspace_bytes);
  NNStats* stats = A.take<NNStats>(1);
  NNGrid* grid = A.take<NNGrid>(1);
  NNCloud c[2];
  char* scan_w[2];
  char* scan_c[2];
  const float* pts[2] = {pc0, pc1};
  const int ns[2] = {n0, n1};
  for (int k = 0; k < 2; ++k) {
    c[k].pts = pts[k];
    c[k].n = ns[k];
    c[k].keys = A.take<unsigned>(ns[k]);
    c[k].bitmap = A.take<unsigned long long>(kNNMaxWords);
    c[k].word_prefix = A.take<int>(kNNMaxWords);
    c[k].count = A.take<int>((size_t)ns[k] + 1);
    c[k].slot = A.take<int>(ns[k]);
    c[k].cell_start = A.take<int>((size_t)ns[k] + 2);
    c[k].sorted = A.take<float4>(ns[k]);
    c[k].n_cells_occ = A.take<int>(1);
    scan_w[k] = A.take<char>(ScanScratch::bytes(kNNMaxWords));
    scan_c[k] = A.take<char>(ScanScratch::bytes((long long)ns[k] + 2));
  }
  if (!A.ok()) return HIMO_ERR_WORKSPACE;

  k_nn_stats_init<<<1, 32, 0, stream>>>(stats);
  HIMO_LAUNCH_RET();
  k_nn_stats<<<min(ceil_div(n0 + n1, 256), kNumSMs * 2), 256, 0, stream>>>(pc0, n0, pc1, n1, stats);
  HIMO_LAUNCH_RET();
  k_nn_params<<<1, 32, 0, stream>>>(stats, cell_size, nn_max_bits((long long)n0 + n1), grid);
  HIMO_LAUNCH_RET();
  const int nmax = n0 > n1 ? n0 : n1;
  dim3 grid2(min(ceil_div(nmax, 256), kNumSMs * 8), 2);
  k_nn_clear<<<dim3(kNumSMs * 4, 2), 256, 0, stream>>>(c[0], c[1], grid);   // n_words is device-side
  HIMO_LAUNCH_RET();
  k_nn_mark<<<grid2, 256, 0, stream>>>(c[0], c[1], grid);
  HIMO_LAUNCH_RET();
  for (int k = 0; k < 2; ++k)
    HIMO_CUDA_RET(scan_exclusive(MapPopc64{c[k].bitmap}, c[k].word_prefix, (int)kNNMaxWords,
                                 &grid->n_words, c[k].n_cells_occ, scan_w[k], stream));
  k_nn_rank<<<grid2, 256, 0, stream>>>(c[0], c[1]);
  HIMO_LAUNCH_RET();
  for (int k = 0; k < 2; ++k)
    HIMO_CUDA_RET(scan_exclusive(MapLoadInt{c[k].count}, c[k].cell_start, ns[k] + 1, nullptr, nullptr,
                                 scan_c[k], stream));
  k_nn_fill<<<grid2, 256, 0, stream>>>(c[0], c[1]);
  HIMO_LAUNCH_RET();
  if (g_nn_warp) {   // one warp per query, 32 resident warps per SM
    dim3 grid3(min(ceil_div(nmax, kNNWWarps), kNumSMs * 16), 2);
    k_nn_search_warp<<<grid3, 32 * kNNWWarps, 0, stream>>>(c[0], c[1], grid, dist0, idx0, dist1, idx1, radius2);
  } else {
    dim3 grid3(ceil_div(nmax, 128), 2);
    k_nn_search<<<grid3, 128, 0, stream>>>(c[0], c[1], grid, dist0, idx0, dist1, idx1, radius2);
  }
  HIMO_LAUNCH_RET();
  return HIMO_OK;
}

extern "C" int himo_chamfer_forward(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                                    float* dist1, int32_t* idx0, int32_t* idx1, float cell_size,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  return nn_forward_impl(pc0, n0, pc1, n1, dist0, dist1, idx0, idx1, cell_size, INFINITY, workspace,
                         workspace_bytes, stream);
}

extern "C" int himo_chamfer_forward_radius(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                                           float* dist1, int32_t* idx0, int32_t* idx1, float radius,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  if (!(radius > 0.f)) return HIMO_ERR_ARG;
  // candidates are accepted on d < best: start just above r^2 so that d == r^2 still counts
  const float r2 = nextafterf(radius * radius, INFINITY);
  return nn_forward_impl(pc0, n0, pc1, n1, dist0, dist1, idx0, idx1, 0.f, r2, workspace,
                         workspace_bytes, stream);
}

extern "C" int himo_chamfer_backward(const float* pc0, int n0, const float* pc1, int n1,
                                     const int32_t* idx0, const int32_t* idx1,
                                     const float* grad_dist0, const float* grad_dist1,
                                     float* grad_pc0, float* grad_pc1, void* stream_) {
  if (n0 < 0 || n1 < 0) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n0 > 0 && n1 > 0) {
    k_chamfer_grad<<<min(ceil_div(n0, 256), kNumSMs * 8), 256, 0, stream>>>(pc0, n0, pc1, idx0, grad_dist0,
                                                                            grad_pc0, grad_pc1);
    HIMO_LAUNCH_RET();
    k_chamfer_grad<<<min(ceil_div(n1, 256), kNumSMs * 8), 256, 0, stream>>>(pc1, n1, pc0, idx1, grad_dist1,
                                                                            grad_pc1, grad_pc0);
    HIMO_LAUNCH_RET();
  }
  return HIMO_OK;
}
