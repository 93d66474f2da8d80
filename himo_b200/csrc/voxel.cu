// himo_b200/csrc/voxel.cu -- H1: dynamic voxelization + point-to-voxel scatter (operator form).
//
// Drop-in replacements, at the C ABI, for the two `mmcv` extension entry points on the hot path:
//   dynamic_voxelize_forward        OSF/assets/cuda/mmcv/pybind.cpp:47-49, voxelization_cuda.cu:246-286
//   dynamic_point_to_voxel_forward  OSF/assets/cuda/mmcv/pybind.cpp:33-35, scatter_points_cuda.cu:9-66
// Not a translation: the reference sorts all N coordinate rows (at::unique_dim) and then issues C
// scalar atomics per point; here the voxel order comes from a one-bit-per-cell occupancy bitmap and a
// popcount scan (O(N + cells/32)), points are counting-sorted into voxel segments and each voxel is
// reduced by one warp with double accumulation (no float atomics => results independent of thread
// scheduling up to a final-rounding tie).
#include "common.cuh"
#include "himo_b200.h"

namespace himo {

// ------------------------------------------------------------------------------ voxelize
struct VoxelGrid {
  float vx, vy, vz;
  float x_min, y_min, z_min;
  int gx, gy, gz;
};

__host__ inline VoxelGrid make_voxel_grid(const float* voxel_size, const float* range) {
  VoxelGrid g;
  g.vx = voxel_size[0]; g.vy = voxel_size[1]; g.vz = voxel_size[2];
  g.x_min = range[0]; g.y_min = range[1]; g.z_min = range[2];
  // fp32 subtraction and division, then round-half-away (voxelization_cuda.cu:269-271)
  g.gx = (int)roundf((range[3] - range[0]) / voxel_size[0]);
  g.gy = (int)roundf((range[4] - range[1]) / voxel_size[1]);
  g.gz = (int)roundf((range[5] - range[2]) / voxel_size[2]);
  return g;
}

// Per-point cell computation: one fp32 subtract, one IEEE divide (__fdiv_rn, never a reciprocal
// multiply), floor, saturating convert -- the exact sequence of voxelization_cuda_kernel.cuh:26-45.
// Returns 0 when inside, 1/2/3 for the first axis (x/y/z) that fails; writes cx,cy,cz.
__device__ __forceinline__ int voxel_cell(const VoxelGrid& g, float x, float y, float z, int& cx,
                                          int& cy, int& cz) {
  cx = __float2int_rd(__fdiv_rn(x - g.x_min, g.vx));
  if (cx < 0 || cx >= g.gx) return 1;
  cy = __float2int_rd(__fdiv_rn(y - g.y_min, g.vy));
  if (cy < 0 || cy >= g.gy) return 2;
  cz = __float2int_rd(__fdiv_rn(z - g.z_min, g.vz));
  if (cz < 0 || cz >= g.gz) return 3;
  return 0;
}

__global__ void __launch_bounds__(256)
k_dynamic_voxelize(const float* __restrict__ points, int n, int nf, VoxelGrid g,
                   int32_t* __restrict__ coors) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float* p = points + (size_t)i * nf;
    int32_t* c = coors + (size_t)i * 3;
    int cx, cy, cz;
    int fail = voxel_cell(g, __ldg(p), __ldg(p + 1), __ldg(p + 2), cx, cy, cz);
    // The reference overwrites only the leading entries on its early exits and relies on the
    // caller's pre-zeroed buffer (voxelize.py:78, its only call site).  We write the full row the
    // reference ends up with under that contract: (-1,0,0), (-1,-1,0) or (-1,-1,-1).
    if (fail == 0) { c[0] = cz; c[1] = cy; c[2] = cx; }
    else { c[0] = -1; c[1] = fail >= 2 ? -1 : 0; c[2] = fail == 3 ? -1 : 0; }
  }
}

// Fast path for tightly packed xyz rows (nf == 3, 16-byte aligned): each thread handles 8 points per
// iteration -- six independent 128-bit streaming loads in flight, then six 128-bit streaming stores
// (96 B in / 96 B out, fully coalesced; neither array is re-read, so both bypass L1 / are evict-first).
__device__ __forceinline__ void stg_stream_i4(int4* p, int4 v) {
  asm volatile("st.global.cs.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void voxel_quad(const VoxelGrid& g, float4 a, float4 b, float4 c, int4* __restrict__ out) {
  const float px[4] = {a.x, a.w, b.z, c.y};
  const float py[4] = {a.y, b.x, b.w, c.z};
  const float pz[4] = {a.z, b.y, c.x, c.w};
  int o[12];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int cx, cy, cz;
    int fail = voxel_cell(g, px[k], py[k], pz[k], cx, cy, cz);
    if (fail == 0) { o[3 * k] = cz; o[3 * k + 1] = cy; o[3 * k + 2] = cx; }
    else { o[3 * k] = -1; o[3 * k + 1] = fail >= 2 ? -1 : 0; o[3 * k + 2] = fail == 3 ? -1 : 0; }
  }
  stg_stream_i4(out, make_int4(o[0], o[1], o[2], o[3]));
  stg_stream_i4(out + 1, make_int4(o[4], o[5], o[6], o[7]));
  stg_stream_i4(out + 2, make_int4(o[8], o[9], o[10], o[11]));
}
__global__ void __launch_bounds__(256)
k_dynamic_voxelize_x4(const float4* __restrict__ points4, int n_quads, VoxelGrid g,
                      int4* __restrict__ coors4) {
  const int stride = gridDim.x * blockDim.x;
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  for (; q + stride < n_quads; q += 2 * stride) {      // two quads (8 points) per iteration
    const int q2 = q + stride;
    const float4 a0 = ldg_stream_f4(points4 + 3 * (size_t)q), b0 = ldg_stream_f4(points4 + 3 * (size_t)q + 1),
                 c0 = ldg_stream_f4(points4 + 3 * (size_t)q + 2);
    const float4 a1 = ldg_stream_f4(points4 + 3 * (size_t)q2), b1 = ldg_stream_f4(points4 + 3 * (size_t)q2 + 1),
                 c1 = ldg_stream_f4(points4 + 3 * (size_t)q2 + 2);
    voxel_quad(g, a0, b0, c0, coors4 + 3 * (size_t)q);
    voxel_quad(g, a1, b1, c1, coors4 + 3 * (size_t)q2);
  }
  if (q < n_quads) {
    const float4 a = ldg_stream_f4(points4 + 3 * (size_t)q), b = ldg_stream_f4(points4 + 3 * (size_t)q + 1),
                 c = ldg_stream_f4(points4 + 3 * (size_t)q + 2);
    voxel_quad(g, a, b, c, coors4 + 3 * (size_t)q);
  }
}

// ------------------------------------------------------------------------------ scatter
template <typename CoorT>
__global__ void __launch_bounds__(256)
k_scatter_mark(const CoorT* __restrict__ coors, int n, int d0, int d1, int d2,
               long long* __restrict__ keys, unsigned* __restrict__ bitmap, int* __restrict__ err) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long a = (long long)coors[3 * (size_t)i], b = (long long)coors[3 * (size_t)i + 1],
              c = (long long)coors[3 * (size_t)i + 2];
    long long key = -1;
    if (a >= 0 && b >= 0 && c >= 0) {
      if (a >= d0 || b >= d1 || c >= d2) { atomicOr(err, 1); }
      else {
        key = (a * d1 + b) * d2 + c;
        atomicOr(bitmap + (key >> 5), 1u << (key & 31));
      }
    }
    keys[i] = key;
  }
}

template <typename CoorT>
__global__ void __launch_bounds__(256)
k_scatter_rank(const long long* __restrict__ keys, int n, int d1, int d2,
               const unsigned* __restrict__ bitmap, const int* __restrict__ word_prefix,
               int32_t* __restrict__ point2voxel, int32_t* __restrict__ slot,
               int32_t* __restrict__ voxel_count, CoorT* __restrict__ voxel_coors) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    long long key = keys[i];
    int rank = -1;
    if (key >= 0) {
      rank = bitmap_rank_lb(bitmap, word_prefix, key);
      int s = atomicAdd(voxel_count + rank, 1);
      slot[i] = s;
      if (s == 0) {
        long long c = key % d2, t = key / d2;
        voxel_coors[3 * (size_t)rank] = (CoorT)(t / d1);
        voxel_coors[3 * (size_t)rank + 1] = (CoorT)(t % d1);
        voxel_coors[3 * (size_t)rank + 2] = (CoorT)c;
      }
    }
    point2voxel[i] = rank;
  }
}

__global__ void __launch_bounds__(256)
k_scatter_fill(const int32_t* __restrict__ point2voxel, const int32_t* __restrict__ slot, int n,
               const int* __restrict__ seg_start, int32_t* __restrict__ sorted_idx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int v = point2voxel[i];
    if (v >= 0) sorted_idx[seg_start[v] + slot[i]] = i;
  }
}

// One warp per voxel, lane = feature channel (channels tiled by 32).  Every lane walks the
// voxel's point segment, so a point's feature row is one coalesced 4*C-byte read.
__global__ void __launch_bounds__(256)
k_scatter_reduce(const float* __restrict__ feats, int c, int reduce_type,
                 const int* __restrict__ num_voxels, const int* __restrict__ seg_start,
                 const int32_t* __restrict__ voxel_count, const int32_t* __restrict__ sorted_idx,
                 float* __restrict__ voxel_feats) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int m = *num_voxels;
  for (int v = blockIdx.x * warps_per_block + (threadIdx.x >> 5); v < m;
       v += gridDim.x * warps_per_block) {
    const int beg = seg_start[v], cnt = voxel_count[v];
    for (int ch = lane; ch < c; ch += 32) {
      if (reduce_type == 2) {
        float best = -INFINITY;
        for (int k = 0; k < cnt; ++k) {
          float f = __ldg(feats + (size_t)sorted_idx[beg + k] * c + ch);
          best = f > best ? f : best;
        }
        voxel_feats[(size_t)v * c + ch] = best;
      } else {
        double acc = 0.0;
        for (int k = 0; k < cnt; ++k) acc += (double)__ldg(feats + (size_t)sorted_idx[beg + k] * c + ch);
        float s = (float)acc;
        // mean: sum rounded to fp32, then an fp32 divide by the fp32 count (scatter_points_cuda.cu:59-60)
        voxel_feats[(size_t)v * c + ch] = reduce_type == 1 ? __fdiv_rn(s, (float)cnt) : s;
      }
    }
  }
}

struct ScatterLayout {
  long long n_cells;
  long long n_words;  // ceil(n_cells/32) + 1 (one zero word so rank(n_cells) is defined)
};

}  // namespace himo

using namespace himo;

extern "C" int himo_dynamic_voxelize_forward(const float* points, int num_points, int num_features,
                                             const float* voxel_size, const float* coors_range,
                                             int32_t* coors, void* stream_) {
  if (num_points < 0 || num_features < 3 || !voxel_size || !coors_range) return HIMO_ERR_ARG;
  if (num_points == 0) return HIMO_OK;
  if (!points || !coors) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  VoxelGrid g = make_voxel_grid(voxel_size, coors_range);
  const bool vec = num_features == 3 && ((uintptr_t)points % 16 == 0) && ((uintptr_t)coors % 16 == 0);
  int done = 0;
  if (vec && num_points >= 4) {
    int quads = num_points / 4;
    int blocks = min(ceil_div(quads, 256), kNumSMs * 8);
    k_dynamic_voxelize_x4<<<blocks, 256, 0, stream>>>((const float4*)points, quads, g, (int4*)coors);
    HIMO_LAUNCH_RET();
    done = quads * 4;
  }
  if (done < num_points) {
    int rest = num_points - done;
    int blocks = min(ceil_div(rest, 256), kNumSMs * 8);
    k_dynamic_voxelize<<<blocks, 256, 0, stream>>>(points + (size_t)done * num_features, rest,
                                                   num_features, g, coors + (size_t)done * 3);
    HIMO_LAUNCH_RET();
  }
  return HIMO_OK;
}

static inline bool scatter_layout(const int32_t* dims, ScatterLayout* L) {
  if (!dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return false;
  long long cells = (long long)dims[0] * dims[1] * dims[2];
  if (cells > (1ll << 33)) return false;
  L->n_cells = cells;
  L->n_words = (cells + 31) / 32 + 1;
  return true;
}

extern "C" size_t himo_dynamic_point_to_voxel_workspace_bytes(int num_points, int num_feats,
                                                              const int32_t* dims) {
  (void)num_feats;
  ScatterLayout L;
  if (num_points < 0 || !scatter_layout(dims, &L)) return 0;
  size_t n = (size_t)(num_points > 0 ? num_points : 1);
  size_t b = 0;
  b += align_up(n * sizeof(long long), 256);               // keys
  b += align_up((size_t)L.n_words * sizeof(unsigned), 256);  // bitmap
  b += align_up((size_t)L.n_words * sizeof(int), 256);       // word_prefix
  b += align_up(n * sizeof(int32_t), 256);                  // slot
  b += align_up((n + 1) * sizeof(int), 256);                // seg_start
  b += align_up(n * sizeof(int32_t), 256);                  // sorted_idx
  b += ScanScratch::bytes(L.n_words) + ScanScratch::bytes((long long)n + 1);
  b += 256;                                                 // err flag
  return b + 4096;
}

template <typename CoorT>
static int scatter_impl(const float* feats, const CoorT* coors, int n, int c, int reduce_type,
                        const int32_t* dims, float* voxel_feats, CoorT* voxel_coors,
                        int32_t* point2voxel, int32_t* voxel_count, int32_t* num_voxels,
                        void* workspace, size_t ws_bytes, cudaStream_t stream) {
  ScatterLayout L;
  if (!scatter_layout(dims, &L)) return HIMO_ERR_ARG;
  if (L.n_words > 0x7fffffffll) return HIMO_ERR_UNSUPPORTED;
  Arena A(workspace, ws_bytes);
  long long* keys = A.take<long long>(n);
  unsigned* bitmap = A.take<unsigned>(L.n_words);
  int* word_prefix = A.take<int>(L.n_words);
  int32_t* slot = A.take<int32_t>(n);
  int* seg_start = A.take<int>((size_t)n + 1);
  int32_t* sorted_idx = A.take<int32_t>(n);
  char* scan1 = A.take<char>(ScanScratch::bytes(L.n_words));
  char* scan2 = A.take<char>(ScanScratch::bytes((long long)n + 1));
  int* err = A.take<int>(1);
  if (!A.ok()) return HIMO_ERR_WORKSPACE;

  HIMO_CUDA_RET(cudaMemsetAsync(bitmap, 0, (size_t)L.n_words * sizeof(unsigned), stream));
  HIMO_CUDA_RET(cudaMemsetAsync(voxel_count, 0, (size_t)n * sizeof(int32_t), stream));
  HIMO_CUDA_RET(cudaMemsetAsync(err, 0, sizeof(int), stream));
  const int blocks = min(ceil_div(n, 256), kNumSMs * 8);
  k_scatter_mark<CoorT><<<blocks, 256, 0, stream>>>(coors, n, dims[0], dims[1], dims[2], keys, bitmap, err);
  HIMO_LAUNCH_RET();
  HIMO_CUDA_RET(scan_exclusive(MapPopc{bitmap}, word_prefix, (int)L.n_words, nullptr, num_voxels, scan1, stream));
  k_scatter_rank<CoorT><<<blocks, 256, 0, stream>>>(keys, n, dims[1], dims[2], bitmap, word_prefix,
                                                    point2voxel, slot, voxel_count, voxel_coors);
  HIMO_LAUNCH_RET();
  // segment starts = exclusive scan of the per-voxel counts (entries >= M are zero)
  HIMO_CUDA_RET(scan_exclusive(MapLoadInt{voxel_count}, seg_start, n, nullptr, nullptr, scan2, stream));
  k_scatter_fill<<<blocks, 256, 0, stream>>>(point2voxel, slot, n, seg_start, sorted_idx);
  HIMO_LAUNCH_RET();
  k_scatter_reduce<<<kNumSMs * 4, 256, 0, stream>>>(feats, c, reduce_type, num_voxels, seg_start,
                                                    voxel_count, sorted_idx, voxel_feats);
  HIMO_LAUNCH_RET();
  // out-of-range coordinate => dims were wrong; report synchronously only if the caller asks
  // (himo_dynamic_point_to_voxel_check); the flag lives at the end of the workspace.
  return HIMO_OK;
}

extern "C" int himo_dynamic_point_to_voxel_forward(
    const float* feats, const void* coors, int coors_is_int64, int num_points, int num_feats,
    int reduce_type, const int32_t* dims, float* voxel_feats, void* voxel_coors,
    int32_t* point2voxel, int32_t* voxel_count, int32_t* num_voxels, void* workspace,
    size_t workspace_bytes, void* stream_) {
  if (num_points < 0 || num_feats <= 0 || reduce_type < 0 || reduce_type > 2) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!num_voxels) return HIMO_ERR_ARG;
  if (num_points == 0) {
    HIMO_CUDA_RET(cudaMemsetAsync(num_voxels, 0, sizeof(int32_t), stream));
    return HIMO_OK;
  }
  if (!feats || !coors || !voxel_feats || !voxel_coors || !point2voxel || !voxel_count)
    return HIMO_ERR_ARG;
  if (coors_is_int64)
    return scatter_impl<long long>(feats, (const long long*)coors, num_points, num_feats, reduce_type,
                                   dims, voxel_feats, (long long*)voxel_coors, point2voxel,
                                   voxel_count, num_voxels, workspace, workspace_bytes, stream);
  return scatter_impl<int32_t>(feats, (const int32_t*)coors, num_points, num_feats, reduce_type, dims,
                               voxel_feats, (int32_t*)voxel_coors, point2voxel, voxel_count,
                               num_voxels, workspace, workspace_bytes, stream);
}
