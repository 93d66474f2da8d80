// himo_b200/csrc/nn.cu -- H2: exact bidirectional 1-NN (Chamfer correspondence) on a uniform cell grid.
//
// Drop-in, at the C ABI, for chamfer3D.forward / chamfer3D.backward
// (OSF/assets/cuda/chamfer3D/chamfer3D_cuda.cpp:18-35, chamfer3D.cu:33-154).
// The reference streams the whole other cloud past every query (O(N0*N1), 256-point smem tiles).
// Here both clouds are counting-sorted into a shared uniform grid (one occupancy bit per cell +
// popcount ranks => CSR cell ranges, no hash collisions to resolve) and every query expands
// Chebyshev rings of cells until the ring bound proves the current best is the global one.  The
// result is the *exact* nearest neighbour with the reference's tie rule (lowest index), and the
// squared distance is evaluated with the reference's rounding sequence fma(dz,dz,fma(dy,dy,dx*dx)).
#include "common.cuh"
#include "himo_b200.h"

namespace himo {

constexpr long long kNNMaxCells = 1ll << 24;              // occupancy bitmap: 2 MiB per cloud
constexpr long long kNNMaxWords = kNNMaxCells / 32 + 1;

struct NNGrid {
  float ox, oy, oz;   // grid origin (bbox min of both clouds)
  float h;            // cell edge
  int nx, ny, nz;
  int n_words;        // ceil(nx*ny*nz/32) + 1
};

// monotone float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_nn_bbox_init(unsigned* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256)
k_nn_bbox(const float* __restrict__ pc0, int n0, const float* __restrict__ pc1, int n1,
          unsigned* __restrict__ bbox) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  const int n = n0 + n1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = i < n0 ? pc0 + 3 * (size_t)i : pc1 + 3 * (size_t)(i - n0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = __ldg(p + k);
      lo[k] = fminf(lo[k], v);
      hi[k] = fmaxf(hi[k], v);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (lo[k] <= hi[k]) {
        atomicMin(bbox + k, f2ord(lo[k]));
        atomicMax(bbox + 3 + k, f2ord(hi[k]));
      }
    }
  }
}

__global__ void k_nn_params(const unsigned* __restrict__ bbox, float cell, NNGrid* __restrict__ g) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo[3], hi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = ord2f(bbox[k]); hi[k] = ord2f(bbox[3 + k]); }
  for (int k = 0; k < 3; ++k)
    if (!(lo[k] <= hi[k]) || !isfinite(lo[k]) || !isfinite(hi[k])) { lo[k] = 0.f; hi[k] = 0.f; }
  float h = cell;
  int nx, ny, nz;
  for (int it = 0; it < 64; ++it) {
    nx = (int)fminf(floorf((hi[0] - lo[0]) / h) + 1.f, 2.0e9f);
    ny = (int)fminf(floorf((hi[1] - lo[1]) / h) + 1.f, 2.0e9f);
    nz = (int)fminf(floorf((hi[2] - lo[2]) / h) + 1.f, 2.0e9f);
    double cells = (double)nx * (double)ny * (double)nz;
    if (cells <= (double)kNNMaxCells) break;
    h *= 1.26f;  // ~ cube root of 2: halve the cell count per step
  }
  g->ox = lo[0]; g->oy = lo[1]; g->oz = lo[2];
  g->h = h;
  g->nx = nx; g->ny = ny; g->nz = nz;
  long long cells = (long long)nx * ny * nz;
  g->n_words = (int)((cells + 31) / 32 + 1);
}

__device__ __forceinline__ void nn_cell(const NNGrid& g, float x, float y, float z, int& cx, int& cy,
                                        int& cz) {
  cx = min(max(__float2int_rd(__fdiv_rn(x - g.ox, g.h)), 0), g.nx - 1);
  cy = min(max(__float2int_rd(__fdiv_rn(y - g.oy, g.h)), 0), g.ny - 1);
  cz = min(max(__float2int_rd(__fdiv_rn(z - g.oz, g.h)), 0), g.nz - 1);
}

struct NNCloud {
  const float* pts;      // [n,3]
  int n;
  int* keys;             // [n]
  unsigned* bitmap;      // [kNNMaxWords]
  int* word_prefix;      // [kNNMaxWords]
  int* count;            // [n+1] zeroed
  int* slot;             // [n]
  int* cell_start;       // [n+2]
  float4* sorted;        // [n] xyz + original index bits
  int* n_cells_occ;      // [1]
};

__global__ void __launch_bounds__(256)
k_nn_mark(NNCloud c0, NNCloud c1, const NNGrid* __restrict__ gp) {
  const NNGrid g = *gp;
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    int cx, cy, cz;
    nn_cell(g, __ldg(c.pts + 3 * (size_t)i), __ldg(c.pts + 3 * (size_t)i + 1),
            __ldg(c.pts + 3 * (size_t)i + 2), cx, cy, cz);
    int key = (cz * g.ny + cy) * g.nx + cx;
    c.keys[i] = key;
    atomicOr(c.bitmap + (key >> 5), 1u << (key & 31));
  }
}

__global__ void __launch_bounds__(256)
k_nn_rank(NNCloud c0, NNCloud c1) {
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    int key = c.keys[i];
    int rank = bitmap_rank_lb(c.bitmap, c.word_prefix, key);
    c.keys[i] = rank;  // keys now hold the occupied-cell rank
    c.slot[i] = atomicAdd(c.count + rank, 1);
  }
}

__global__ void __launch_bounds__(256)
k_nn_fill(NNCloud c0, NNCloud c1) {
  const NNCloud& c = blockIdx.y == 0 ? c0 : c1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x) {
    int pos = c.cell_start[c.keys[i]] + c.slot[i];
    c.sorted[pos] = make_float4(__ldg(c.pts + 3 * (size_t)i), __ldg(c.pts + 3 * (size_t)i + 1),
                                __ldg(c.pts + 3 * (size_t)i + 2), __int_as_float(i));
  }
}

// One thread per query, queries taken in cell-sorted order so that the lanes of a warp walk the
// same reference ranges (L1 broadcast).  Both directions run in one launch (blockIdx.y).
__global__ void __launch_bounds__(128)
k_nn_search(NNCloud c0, NNCloud c1, const NNGrid* __restrict__ gp, float* __restrict__ dist0,
            int32_t* __restrict__ idx0, float* __restrict__ dist1, int32_t* __restrict__ idx1) {
  const NNGrid g = *gp;
  const NNCloud& q = blockIdx.y == 0 ? c0 : c1;
  const NNCloud& r = blockIdx.y == 0 ? c1 : c0;
  float* __restrict__ dist = blockIdx.y == 0 ? dist0 : dist1;
  int32_t* __restrict__ idx = blockIdx.y == 0 ? idx0 : idx1;
  const unsigned* __restrict__ bitmap = r.bitmap;
  const int* __restrict__ prefix = r.word_prefix;
  const int* __restrict__ cstart = r.cell_start;
  const float4* __restrict__ rs = r.sorted;
  const int max_ring = max(g.nx, max(g.ny, g.nz));
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < q.n; t += gridDim.x * blockDim.x) {
    const float4 p = q.sorted[t];
    int cx, cy, cz;
    nn_cell(g, p.x, p.y, p.z, cx, cy, cz);
    // distance from the query to the nearest face of its own cell (0 if it was clamped in)
    float fx = p.x - (g.ox + (float)cx * g.h), fy = p.y - (g.oy + (float)cy * g.h),
          fz = p.z - (g.oz + (float)cz * g.h);
    float dface = fminf(fminf(fminf(fx, g.h - fx), fminf(fy, g.h - fy)), fminf(fz, g.h - fz));
    dface = fmaxf(dface, 0.f);
    float best = 1e20f;
    int best_i = -1;
    for (int ring = 0; ring <= max_ring; ++ring) {
      const int z_lo = max(cz - ring, 0), z_hi = min(cz + ring, g.nz - 1);
      const int y_lo = max(cy - ring, 0), y_hi = min(cy + ring, g.ny - 1);
      for (int z = z_lo; z <= z_hi; ++z) {
        const bool z_shell = (z == cz - ring) || (z == cz + ring);
        for (int y = y_lo; y <= y_hi; ++y) {
          const bool shell = z_shell || (y == cy - ring) || (y == cy + ring);
          // on the shell take the whole x-row, inside only its two end cells
          const int nseg = (shell || ring == 0) ? 1 : 2;
          for (int s = 0; s < nseg; ++s) {
            int xa, xb;
            if (nseg == 1) { xa = cx - ring; xb = cx + ring; }
            else { xa = xb = (s == 0 ? cx - ring : cx + ring); }
            if (xb < 0 || xa >= g.nx) continue;
            xa = max(xa, 0); xb = min(xb, g.nx - 1);
            const int row = (z * g.ny + y) * g.nx;
            const int beg = cstart[bitmap_rank_lb(bitmap, prefix, row + xa)];
            const int end = cstart[bitmap_rank_lb(bitmap, prefix, row + xb + 1)];
            for (int j = beg; j < end; ++j) {
              const float4 c = rs[j];
              const float dx = c.x - p.x, dy = c.y - p.y, dz = c.z - p.z;
              const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
              const int ci = __float_as_int(c.w);
              if (d < best || (d == best && ci < best_i)) { best = d; best_i = ci; }
            }
          }
        }
      }
      // every unexamined point lies outside ring `ring`: at least ring*h + dface away
      // (small slack absorbs fp32 rounding in the cell assignment)
      const float bound = (float)ring * g.h + dface - 1e-3f * g.h;
      if (best_i >= 0 && bound > 0.f && best < bound * bound) break;
      if (cx - ring <= 0 && cx + ring >= g.nx - 1 && cy - ring <= 0 && cy + ring >= g.ny - 1 &&
          cz - ring <= 0 && cz + ring >= g.nz - 1)
        break;  // whole grid examined
    }
    const int qi = __float_as_int(p.w);
    dist[qi] = best;
    idx[qi] = best_i;
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

static size_t nn_cloud_bytes(int n) {
  size_t m = (size_t)(n > 0 ? n : 1);
  size_t b = 0;
  b += align_up(m * sizeof(int), 256);                    // keys
  b += align_up((size_t)kNNMaxWords * 4, 256) * 2;         // bitmap + prefix
  b += align_up((m + 1) * sizeof(int), 256);              // count
  b += align_up(m * sizeof(int), 256);                    // slot
  b += align_up((m + 2) * sizeof(int), 256);              // cell_start
  b += align_up(m * sizeof(float4), 256);                 // sorted
  b += 256;                                               // n_cells_occ
  b += ScanScratch::bytes(kNNMaxWords) + ScanScratch::bytes((long long)m + 2);
  return b;
}

}  // namespace himo

using namespace himo;

extern "C" size_t himo_chamfer_workspace_bytes(int n0, int n1) {
  if (n0 < 0 || n1 < 0) return 0;
  return nn_cloud_bytes(n0) + nn_cloud_bytes(n1) + 4096 + 16 * 256;
}

extern "C" int himo_chamfer_forward(const float* pc0, int n0, const float* pc1, int n1, float* dist0,
                                    float* dist1, int32_t* idx0, int32_t* idx1, float cell_size,
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
  if (!(cell_size > 0.f)) cell_size = 0.5f;
  Arena A(workspace, workspace_bytes);
  unsigned* bbox = A.take<unsigned>(8);
  NNGrid* grid = A.take<NNGrid>(1);
  NNCloud c[2];
  char* scan_w[2];
  char* scan_c[2];
  const float* pts[2] = {pc0, pc1};
  const int ns[2] = {n0, n1};
  for (int k = 0; k < 2; ++k) {
    c[k].pts = pts[k];
    c[k].n = ns[k];
    c[k].keys = A.take<int>(ns[k]);
    c[k].bitmap = A.take<unsigned>(kNNMaxWords);
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

  k_nn_bbox_init<<<1, 32, 0, stream>>>(bbox);
  HIMO_LAUNCH_RET();
  k_nn_bbox<<<min(ceil_div(n0 + n1, 256), kNumSMs * 4), 256, 0, stream>>>(pc0, n0, pc1, n1, bbox);
  HIMO_LAUNCH_RET();
  k_nn_params<<<1, 32, 0, stream>>>(bbox, cell_size, grid);
  HIMO_LAUNCH_RET();
  for (int k = 0; k < 2; ++k) {
    // the bitmap only needs clearing up to n_words, which is device-side; clear the cap (2 MiB)
    HIMO_CUDA_RET(cudaMemsetAsync(c[k].bitmap, 0, (size_t)kNNMaxWords * 4, stream));
    HIMO_CUDA_RET(cudaMemsetAsync(c[k].count, 0, ((size_t)ns[k] + 1) * sizeof(int), stream));
  }
  const int nmax = n0 > n1 ? n0 : n1;
  dim3 grid2(min(ceil_div(nmax, 256), kNumSMs * 8), 2);
  k_nn_mark<<<grid2, 256, 0, stream>>>(c[0], c[1], grid);
  HIMO_LAUNCH_RET();
  for (int k = 0; k < 2; ++k)
    HIMO_CUDA_RET(scan_exclusive(MapPopc{c[k].bitmap}, c[k].word_prefix, (int)kNNMaxWords,
                                 &grid->n_words, c[k].n_cells_occ, scan_w[k], stream));
  k_nn_rank<<<grid2, 256, 0, stream>>>(c[0], c[1]);
  HIMO_LAUNCH_RET();
  for (int k = 0; k < 2; ++k)
    HIMO_CUDA_RET(scan_exclusive(MapLoadInt{c[k].count}, c[k].cell_start, ns[k] + 1, nullptr, nullptr,
                                 scan_c[k], stream));
  k_nn_fill<<<grid2, 256, 0, stream>>>(c[0], c[1]);
  HIMO_LAUNCH_RET();
  dim3 grid3(ceil_div(nmax, 128), 2);
  k_nn_search<<<grid3, 128, 0, stream>>>(c[0], c[1], grid, dist0, idx0, dist1, idx1);
  HIMO_LAUNCH_RET();
  return HIMO_OK;
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
