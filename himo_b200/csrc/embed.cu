// himo_b200/csrc/embed.cu -- H1+H4 front end: the fused SeFlow++ point embedder.
//
// One call = DynamicEmbedder.forward for the F frames of a frame tuple
// (OSF/src/models/basic/encoder.py:602-631), i.e. per frame:
//   rigid warp into the pc1 frame        wrap_batch_pcs, OSF/src/models/basic/__init__.py:50,57
//   NaN filter + dynamic voxelization    DynamicVoxelizer.forward, encoder.py:567-600
//   cluster mean scatter (C=3)           DynamicPillarFeatureNet.forward, encoder.py:442
//   9-ch decoration, Linear(9,32)+BN+ReLU  encoder.py:439-467, 362-371
//   feature mean scatter (C=32)          encoder.py:468
//   pseudo-image scatter                 PointPillarsScatter.forward_single, encoder.py:126-147
// The reference runs ~40 kernels with >= 3 host syncs per frame here (two radix sorts in
// at::unique_dim, boolean-mask compactions, a dense 3x262144 canvas just to broadcast means).
// This file does it in 6 launches for all frames together, without a host sync: occupancy bitmap +
// popcount ranks give the sorted voxel order, points are counting-sorted into voxel segments, and one
// warp per voxel (lane = feature channel) computes mean -> decoration -> PFN -> mean and writes the
// NHWC split-bf16 canvas row that the tcgen05 convolution reads through TMA.
#include "common.cuh"
#include "himo_b200.h"
#include "umma.cuh"

namespace himo {

struct EmbedGrid {
  float vx, vy, vz, x_min, y_min, z_min;
  int gx, gy, gz;
  float half_vx, half_vy, half_vz;      // voxel_size / 2 in fp32 (encoder.py:519: voxel_size / 2)
  float x_off, y_off, z_off;            // fp32(v/2 + min) formed in double (encoder.py:257-259)
};

struct EmbedFrame {
  const float* pts;     // [n,3] input cloud (sensor frame)
  int n;
  int has_T;
  float T[12];          // row-major 3x4 rigid transform into the target frame
};

struct EmbedArgs {
  EmbedFrame fr[HIMO_MAX_FRAMES];
  int n_frames;
  int n_max;            // row stride of the per-frame arrays
  int n_cells;          // gx*gy
  int n_words;          // n_cells/32 + 1
  float4* pt4;          // [F][n_max]  warped xyz + cell key bits (-1 invalid)
  unsigned* bitmap;     // [F][n_words]
  int* word_prefix;     // [F][n_words]
  int* num_voxels;      // [F]
  int* rank;            // [F][n_max]
  int* slot;            // [F][n_max]
  int* count;           // [F][n_max+1]
  int* seg_start;       // [F][n_max+1]
  int* sorted_idx;      // [F][n_max]
  float* voxel_feats;   // [F][n_max][32] fp32 voxel features in sorted-voxel order
  float* voxel_mean;    // [F][n_max][3] (optional, may be null)
  int* voxel_key;       // [F][n_max] cell key of each voxel (optional, may be null)
  int* big_count;       // [F] number of crowded pillars (> kPfnBigVoxel points), zeroed per call
  int* big_list;        // [F][n_max / kPfnBigVoxel + 1] their voxel indices, in arrival order
  int big_stride;
  __nv_bfloat16* canvas;       // [planes][gy][gx][F*32]
  int canvas_planes;
  long long canvas_plane_stride;
  int canvas_ld;               // channels per pixel of the buffer the canvas lives in (>= F*32)
  const float* pfn_w;   // [32][9] BN-folded
  const float* pfn_b;   // [32]    BN-folded
};

__global__ void __launch_bounds__(256)
k_embed_points(EmbedArgs a, EmbedGrid g) {
  const int f = blockIdx.y;
  const EmbedFrame& fr = a.fr[f];
  float4* pt4 = a.pt4 + (size_t)f * a.n_max;
  unsigned* bitmap = a.bitmap + (size_t)f * a.n_words;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fr.n; i += gridDim.x * blockDim.x) {
    float x = __ldg(fr.pts + 3 * (size_t)i), y = __ldg(fr.pts + 3 * (size_t)i + 1),
          z = __ldg(fr.pts + 3 * (size_t)i + 2);
    if (fr.has_T) {
      // pc @ R^T + t exactly as the fp32 CPU GEMM evaluates it: k-ascending FMA chain, then + t
      const float* T = fr.T;
      float nx = __fadd_rn(__fmaf_rn(z, T[2], __fmaf_rn(y, T[1], __fmul_rn(x, T[0]))), T[3]);
      float ny = __fadd_rn(__fmaf_rn(z, T[6], __fmaf_rn(y, T[5], __fmul_rn(x, T[4]))), T[7]);
      float nz = __fadd_rn(__fmaf_rn(z, T[10], __fmaf_rn(y, T[9], __fmul_rn(x, T[8]))), T[11]);
      x = nx; y = ny; z = nz;
    }
    int key = -1;
    if (!(isnan(x) || isnan(y) || isnan(z))) {      // NaN rows = batch padding (encoder.py:576-578)
      int cx = __float2int_rd(__fdiv_rn(x - g.x_min, g.vx));
      if (cx >= 0 && cx < g.gx) {
        int cy = __float2int_rd(__fdiv_rn(y - g.y_min, g.vy));
        if (cy >= 0 && cy < g.gy) {
          int cz = __float2int_rd(__fdiv_rn(z - g.z_min, g.vz));
          if (cz >= 0 && cz < g.gz) key = cy * g.gx + cx;   // gz == 1 for the pillar encoder
        }
      }
    }
    pt4[i] = make_float4(x, y, z, __int_as_float(key));
    if (key >= 0) atomicOr(bitmap + (key >> 5), 1u << (key & 31));
  }
}

// One block per frame: exclusive popcount scan over the occupancy words.
__global__ void __launch_bounds__(1024)
k_embed_scan_bitmap(EmbedArgs a) {
  __shared__ int s_scan[33];
  __shared__ int s_carry;
  const int f = blockIdx.x;
  const unsigned* bitmap = a.bitmap + (size_t)f * a.n_words;
  int* prefix = a.word_prefix + (size_t)f * a.n_words;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < a.n_words; base += 1024 * 8) {
    int v[8], sum = 0;
    const int i0 = base + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = (i0 + k < a.n_words) ? __popc(bitmap[i0 + k]) : 0; sum += v[k]; }
    int total;
    int excl = block_excl_scan(sum, s_scan, &total);
    int run = s_carry + excl;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (i0 + k < a.n_words) prefix[i0 + k] = run; run += v[k]; }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.num_voxels[f] = s_carry;
}

__global__ void __launch_bounds__(256)
k_embed_rank(EmbedArgs a) {
  const int f = blockIdx.y;
  const int n = a.fr[f].n;
  const float4* pt4 = a.pt4 + (size_t)f * a.n_max;
  const unsigned* bitmap = a.bitmap + (size_t)f * a.n_words;
  const int* prefix = a.word_prefix + (size_t)f * a.n_words;
  int* rank = a.rank + (size_t)f * a.n_max;
  int* slot = a.slot + (size_t)f * a.n_max;
  int* count = a.count + (size_t)f * (a.n_max + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int key = __float_as_int(pt4[i].w);
    int r = -1;
    if (key >= 0) {
      r = bitmap_rank_lb(bitmap, prefix, key);
      slot[i] = atomicAdd(count + r, 1);
    }
    rank[i] = r;
  }
}

// One block per frame: exclusive scan of the per-voxel point counts -> segment starts.
__global__ void __launch_bounds__(1024)
k_embed_scan_counts(EmbedArgs a) {
  __shared__ int s_scan[33];
  __shared__ int s_carry;
  const int f = blockIdx.x;
  const int m = a.num_voxels[f];
  const int* count = a.count + (size_t)f * (a.n_max + 1);
  int* seg = a.seg_start + (size_t)f * (a.n_max + 1);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < m + 1; base += 1024 * 8) {
    int v[8], sum = 0;
    const int i0 = base + threadIdx.x * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = (i0 + k < m) ? count[i0 + k] : 0; sum += v[k]; }
    int total;
    int excl = block_excl_scan(sum, s_scan, &total);
    int run = s_carry + excl;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (i0 + k <= m) seg[i0 + k] = run; run += v[k]; }
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_embed_fill(EmbedArgs a) {
  const int f = blockIdx.y;
  const int n = a.fr[f].n;
  const int* rank = a.rank + (size_t)f * a.n_max;
  const int* slot = a.slot + (size_t)f * a.n_max;
  const int* seg = a.seg_start + (size_t)f * (a.n_max + 1);
  int* sorted_idx = a.sorted_idx + (size_t)f * a.n_max;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = rank[i];
    if (r >= 0) sorted_idx[seg[r] + slot[i]] = i;
  }
}

// Pillars with more points than this are handled by a whole block (k_embed_pfn_big).  Near the sensor a pillar
// holds 500-750 points of a 100 k-point sweep; walked by one warp (4 points per dependent-load round) the largest
// pillar alone set the kernel time (~100 us), although 140 such pillars hold only a third of the points.
constexpr int kPfnBigVoxel = 128;

// One warp per voxel, lane = output feature channel.
__global__ void __launch_bounds__(256)
k_embed_pfn(EmbedArgs a, EmbedGrid g) {
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int m = a.num_voxels[f];
  const float4* pt4 = a.pt4 + (size_t)f * a.n_max;
  const int* seg = a.seg_start + (size_t)f * (a.n_max + 1);
  const int* sorted_idx = a.sorted_idx + (size_t)f * a.n_max;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = __ldg(a.pfn_w + lane * 9 + k);
  const float b = __ldg(a.pfn_b + lane);
  const int C = a.n_frames * 32;
  for (int v = blockIdx.x * warps_per_block + (threadIdx.x >> 5); v < m; v += gridDim.x * warps_per_block) {
    const int beg = seg[v], end = seg[v + 1], cnt = end - beg;
    if (cnt > kPfnBigVoxel) {                    // k_embed_pfn_big: one block per crowded pillar
      if (lane == 0) a.big_list[(size_t)f * a.big_stride + atomicAdd(a.big_count + f, 1)] = v;
      continue;
    }
    // voxel mean of xyz: double accumulation (order-free up to the final rounding), then an fp32
    // divide by the fp32 count as scatter_points_cuda.cu:59-60 does
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int key = 0;
    for (int k = beg + lane; k < end; k += 32) {
      const float4 p = pt4[sorted_idx[k]];
      sx += (double)p.x; sy += (double)p.y; sz += (double)p.z;
      key = __float_as_int(p.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, d);
      sy += __shfl_xor_sync(0xffffffffu, sy, d);
      sz += __shfl_xor_sync(0xffffffffu, sz, d);
    }
    key = __shfl_sync(0xffffffffu, key, 0);      // lane 0 always owns the first point of the segment
    const float fc = (float)cnt;
    const float mx = __fdiv_rn((float)sx, fc), my = __fdiv_rn((float)sy, fc), mz = __fdiv_rn((float)sz, fc);
    const int cy = key / g.gx, cx = key - cy * g.gx;
    // PFN voxel centre: fl(fl(c * v) + fl32(v/2 + min))   (encoder.py:453-458 with :257-259)
    const float ccx = __fadd_rn(__fmul_rn((float)cx, g.vx), g.x_off);
    const float ccy = __fadd_rn(__fmul_rn((float)cy, g.vy), g.y_off);
    const float ccz = __fadd_rn(__fmul_rn(0.f, g.vz), g.z_off);
    // 4 points per trip with independent accumulators: the per-point chain (index -> point -> 9 FMAs ->
    // double add) is latency-bound, and real sweeps have pillars with > 1000 points
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    auto pfn_point = [&](const float4 p) -> double {
      float y = w[0] * p.x;
      y = fmaf(w[1], p.y, y);
      y = fmaf(w[2], p.z, y);
      y = fmaf(w[3], p.x - mx, y);
      y = fmaf(w[4], p.y - my, y);
      y = fmaf(w[5], p.z - mz, y);
      y = fmaf(w[6], p.x - ccx, y);
      y = fmaf(w[7], p.y - ccy, y);
      y = fmaf(w[8], p.z - ccz, y);
      y += b;
      return (double)fmaxf(y, 0.f);
    };
    int k = beg;
    for (; k + 4 <= end; k += 4) {
      // same address for all lanes: each load is one broadcast transaction
      const int i0 = sorted_idx[k], i1 = sorted_idx[k + 1], i2 = sorted_idx[k + 2], i3 = sorted_idx[k + 3];
      const float4 p0 = pt4[i0], p1 = pt4[i1], p2 = pt4[i2], p3 = pt4[i3];
      acc0 += pfn_point(p0); acc1 += pfn_point(p1); acc2 += pfn_point(p2); acc3 += pfn_point(p3);
    }
    for (; k < end; ++k) acc0 += pfn_point(pt4[sorted_idx[k]]);
    const double acc = (acc0 + acc1) + (acc2 + acc3);
    const float feat = __fdiv_rn((float)acc, fc);
    a.voxel_feats[((size_t)f * a.n_max + v) * 32 + lane] = feat;
    if (a.voxel_mean && lane < 3)
      a.voxel_mean[((size_t)f * a.n_max + v) * 3 + lane] = lane == 0 ? mx : (lane == 1 ? my : mz);
    if (a.voxel_key && lane == 0) a.voxel_key[(size_t)f * a.n_max + v] = key;
    // canvas[:, y*512+x] = voxel_feats.T  (encoder.py:140-146), NHWC, channel slice of this frame
    umma::store_split(a.canvas + (size_t)key * a.canvas_ld + f * 32 + lane, a.canvas_plane_stride, a.canvas_planes, feat);
  }
}

// Crowded pillars: one block (8 warps) per voxel.  The mean is a block-wide strided double sum, the PFN sum is
// split into 8 contiguous point ranges (one per warp, lane = channel, 4 independent accumulators) whose partials
// are added in warp order -- fixed association, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256)
k_embed_pfn_big(EmbedArgs a, EmbedGrid g) {
  __shared__ double s_mean[8][3];
  __shared__ double s_part[8][32];
  __shared__ int s_key;
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = a.num_voxels[f];
  const float4* pt4 = a.pt4 + (size_t)f * a.n_max;
  const int* seg = a.seg_start + (size_t)f * (a.n_max + 1);
  const int* sorted_idx = a.sorted_idx + (size_t)f * a.n_max;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) w[k] = __ldg(a.pfn_w + lane * 9 + k);
  const float b = __ldg(a.pfn_b + lane);
  const int C = a.n_frames * 32;
  // the crowded pillars were listed by k_embed_pfn (arrival order; every pillar's result is independent of it)
  const int nbig = a.big_count[f];
  const int* list = a.big_list + (size_t)f * a.big_stride;
  {
  for (int li = blockIdx.x; li < nbig; li += gridDim.x) {
    const int v = list[li];
    const int beg = seg[v], end = seg[v + 1], cnt = end - beg;
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int k = beg + threadIdx.x; k < end; k += 256) {
      const float4 p = pt4[sorted_idx[k]];
      sx += (double)p.x; sy += (double)p.y; sz += (double)p.z;
      if (k == beg) s_key = __float_as_int(p.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, d);
      sy += __shfl_xor_sync(0xffffffffu, sy, d);
      sz += __shfl_xor_sync(0xffffffffu, sz, d);
    }
    if (lane == 0) { s_mean[warp][0] = sx; s_mean[warp][1] = sy; s_mean[warp][2] = sz; }
    __syncthreads();
    sx = 0.0; sy = 0.0; sz = 0.0;
    for (int wv = 0; wv < 8; ++wv) { sx += s_mean[wv][0]; sy += s_mean[wv][1]; sz += s_mean[wv][2]; }
    const int key = s_key;
    const float fc = (float)cnt;
    const float mx = __fdiv_rn((float)sx, fc), my = __fdiv_rn((float)sy, fc), mz = __fdiv_rn((float)sz, fc);
    const int cy = key / g.gx, cx = key - cy * g.gx;
    const float ccx = __fadd_rn(__fmul_rn((float)cx, g.vx), g.x_off);
    const float ccy = __fadd_rn(__fmul_rn((float)cy, g.vy), g.y_off);
    const float ccz = __fadd_rn(__fmul_rn(0.f, g.vz), g.z_off);
    auto pfn_point = [&](const float4 p) -> double {
      float y = w[0] * p.x;
      y = fmaf(w[1], p.y, y);
      y = fmaf(w[2], p.z, y);
      y = fmaf(w[3], p.x - mx, y);
      y = fmaf(w[4], p.y - my, y);
      y = fmaf(w[5], p.z - mz, y);
      y = fmaf(w[6], p.x - ccx, y);
      y = fmaf(w[7], p.y - ccy, y);
      y = fmaf(w[8], p.z - ccz, y);
      y += b;
      return (double)fmaxf(y, 0.f);
    };
    const int per = (cnt + 7) / 8;
    const int wb = beg + warp * per, we = min(wb + per, end);
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int k = wb;
    for (; k + 4 <= we; k += 4) {
      const int i0 = sorted_idx[k], i1 = sorted_idx[k + 1], i2 = sorted_idx[k + 2], i3 = sorted_idx[k + 3];
      const float4 p0 = pt4[i0], p1 = pt4[i1], p2 = pt4[i2], p3 = pt4[i3];
      acc0 += pfn_point(p0); acc1 += pfn_point(p1); acc2 += pfn_point(p2); acc3 += pfn_point(p3);
    }
    for (; k < we; ++k) acc0 += pfn_point(pt4[sorted_idx[k]]);
    s_part[warp][lane] = (acc0 + acc1) + (acc2 + acc3);
    __syncthreads();
    if (warp == 0) {
      double acc = 0.0;
      for (int wv = 0; wv < 8; ++wv) acc += s_part[wv][lane];
      const float feat = __fdiv_rn((float)acc, fc);
      a.voxel_feats[((size_t)f * a.n_max + v) * 32 + lane] = feat;
      if (a.voxel_mean && lane < 3)
        a.voxel_mean[((size_t)f * a.n_max + v) * 3 + lane] = lane == 0 ? mx : (lane == 1 ? my : mz);
      if (a.voxel_key && lane == 0) a.voxel_key[(size_t)f * a.n_max + v] = key;
      umma::store_split(a.canvas + (size_t)key * a.canvas_ld + f * 32 + lane, a.canvas_plane_stride, a.canvas_planes, feat);
    }
    __syncthreads();
  }
  }
}

}  // namespace himo

using namespace himo;

static inline size_t embed_ws_layout(int n_frames, int n_max, int n_words, EmbedArgs* a, void* base) {
  Arena A(base, (size_t)-1);
  const size_t F = (size_t)n_frames, N = (size_t)(n_max > 0 ? n_max : 1);
  float4* pt4 = A.take<float4>(F * N);
  unsigned* bitmap = A.take<unsigned>(F * n_words);
  int* count = A.take<int>(F * (N + 1));          // bitmap and count are contiguous-ish: zeroed together
  int* word_prefix = A.take<int>(F * n_words);
  int* num_voxels = A.take<int>(F);
  int* rank = A.take<int>(F * N);
  int* slot = A.take<int>(F * N);
  int* seg = A.take<int>(F * (N + 1));
  int* sorted_idx = A.take<int>(F * N);
  float* vfeat = A.take<float>(F * N * 32);
  float* vmean = A.take<float>(F * N * 3);
  int* vkey = A.take<int>(F * N);
  const size_t big_stride = N / kPfnBigVoxel + 1;
  int* big_count = A.take<int>(F);
  int* big_list = A.take<int>(F * big_stride);
  if (a) {
    a->pt4 = pt4; a->bitmap = bitmap; a->count = count; a->word_prefix = word_prefix;
    a->num_voxels = num_voxels; a->rank = rank; a->slot = slot; a->seg_start = seg;
    a->sorted_idx = sorted_idx; a->voxel_feats = vfeat; a->voxel_mean = vmean; a->voxel_key = vkey;
    a->big_count = big_count; a->big_list = big_list; a->big_stride = (int)big_stride;
  }
  return A.off + 256;
}

extern "C" size_t himo_embed_workspace_bytes(int n_frames, int n_max, const float* voxel_size,
                                             const float* coors_range) {
  if (n_frames <= 0 || n_frames > HIMO_MAX_FRAMES || n_max < 0 || !voxel_size || !coors_range) return 0;
  const int gx = (int)roundf((coors_range[3] - coors_range[0]) / voxel_size[0]);
  const int gy = (int)roundf((coors_range[4] - coors_range[1]) / voxel_size[1]);
  return embed_ws_layout(n_frames, n_max, gx * gy / 32 + 1, nullptr, nullptr);
}

namespace himo {
// zero the channel slice [0, chunks*16 bytes) of every pixel of an NHWC buffer with `ld16` 16-byte units per pixel
// (cudaMemset2DAsync does this at a fraction of the HBM rate for 192-byte rows)
__global__ void __launch_bounds__(256)
k_clear_slice(uint4* __restrict__ base, int ld16, int chunks, long long total) {
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / chunks;
    const int c = (int)(i - pix * chunks);
    base[pix * ld16 + c] = z;
  }
}
}  // namespace himo

extern "C" int himo_embed_frames(const himo_embed_desc* d, void* stream_) {
  if (!d || d->n_frames <= 0 || d->n_frames > HIMO_MAX_FRAMES || !d->canvas || !d->workspace) return HIMO_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  EmbedGrid g;
  g.vx = d->voxel_size[0]; g.vy = d->voxel_size[1]; g.vz = d->voxel_size[2];
  g.x_min = d->coors_range[0]; g.y_min = d->coors_range[1]; g.z_min = d->coors_range[2];
  g.gx = (int)roundf((d->coors_range[3] - d->coors_range[0]) / d->voxel_size[0]);
  g.gy = (int)roundf((d->coors_range[4] - d->coors_range[1]) / d->voxel_size[1]);
  g.gz = (int)roundf((d->coors_range[5] - d->coors_range[2]) / d->voxel_size[2]);
  if (g.gz != 1 || (g.gx * g.gy) % 32) return HIMO_ERR_UNSUPPORTED;   // pillar encoder: one z bin
  g.half_vx = d->voxel_size[0] / 2; g.half_vy = d->voxel_size[1] / 2; g.half_vz = d->voxel_size[2] / 2;
  // python: self.x_offset = self.vx / 2 + point_cloud_range[0] in double, used as an fp32 scalar
  g.x_off = (float)((double)d->voxel_size_f64[0] / 2 + d->coors_range_f64[0]);
  g.y_off = (float)((double)d->voxel_size_f64[1] / 2 + d->coors_range_f64[1]);
  g.z_off = (float)((double)d->voxel_size_f64[2] / 2 + d->coors_range_f64[2]);

  EmbedArgs a;
  a.n_frames = d->n_frames;
  int n_max = 0;
  for (int f = 0; f < d->n_frames; ++f) {
    a.fr[f].pts = d->points[f];
    a.fr[f].n = d->num_points[f];
    if (a.fr[f].n < 0 || (a.fr[f].n > 0 && !a.fr[f].pts)) return HIMO_ERR_ARG;
    a.fr[f].has_T = d->has_transform[f];
    for (int k = 0; k < 12; ++k) a.fr[f].T[k] = d->transform[f][k];
    n_max = a.fr[f].n > n_max ? a.fr[f].n : n_max;
  }
  if (n_max > d->n_max) return HIMO_ERR_ARG;
  a.n_max = d->n_max;
  a.n_cells = g.gx * g.gy;
  a.n_words = a.n_cells / 32 + 1;
  const size_t need = embed_ws_layout(d->n_frames, d->n_max, a.n_words, &a, d->workspace);
  if (need > d->workspace_bytes) return HIMO_ERR_WORKSPACE;
  // the canvas may be a channel slice [canvas_ch_off, +F*32) of a wider NHWC buffer (the backbone's concatenation
  // buffer: the skip connection of the last UpsampleSkip block then costs no copy)
  a.canvas_ld = d->canvas_ld > 0 ? d->canvas_ld : d->n_frames * 32;
  if (a.canvas_ld < d->n_frames * 32 + d->canvas_ch_off || d->canvas_ch_off < 0 || (a.canvas_ld % 8) || (d->canvas_ch_off % 8))
    return HIMO_ERR_ARG;
  a.canvas = (__nv_bfloat16*)d->canvas + d->canvas_ch_off;
  a.canvas_planes = d->canvas_planes;
  a.canvas_plane_stride = (long long)g.gx * g.gy * a.canvas_ld;
  a.pfn_w = d->pfn_weight; a.pfn_b = d->pfn_bias;
  if (!a.pfn_w || !a.pfn_b || (d->canvas_planes != 1 && d->canvas_planes != 2)) return HIMO_ERR_ARG;

  const size_t F = (size_t)d->n_frames, N = (size_t)(d->n_max > 0 ? d->n_max : 1);
  HIMO_CUDA_RET(cudaMemsetAsync(a.bitmap, 0, F * a.n_words * sizeof(unsigned), stream));
  HIMO_CUDA_RET(cudaMemsetAsync(a.count, 0, F * (N + 1) * sizeof(int), stream));
  HIMO_CUDA_RET(cudaMemsetAsync(a.big_count, 0, F * sizeof(int), stream));
  if (!d->skip_canvas_clear) {
    if (a.canvas_ld == d->n_frames * 32)
      HIMO_CUDA_RET(cudaMemsetAsync(a.canvas, 0, (size_t)a.canvas_plane_stride * d->canvas_planes * 2, stream));
    else {  // channel slice of a wider buffer: the planes follow each other at the same pitch, one strided clear
      const int chunks = d->n_frames * 32 * 2 / 16;
      const long long total = (long long)g.gx * g.gy * d->canvas_planes * chunks;
      k_clear_slice<<<kNumSMs * 8, 256, 0, stream>>>((uint4*)a.canvas, a.canvas_ld * 2 / 16, chunks, total);
      HIMO_LAUNCH_RET();
    }
  }
  if (n_max > 0) {
    dim3 gp(min(ceil_div(n_max, 256), kNumSMs * 8), d->n_frames);
    k_embed_points<<<gp, 256, 0, stream>>>(a, g);
    HIMO_LAUNCH_RET();
    k_embed_scan_bitmap<<<d->n_frames, 1024, 0, stream>>>(a);
    HIMO_LAUNCH_RET();
    k_embed_rank<<<gp, 256, 0, stream>>>(a);
    HIMO_LAUNCH_RET();
    k_embed_scan_counts<<<d->n_frames, 1024, 0, stream>>>(a);
    HIMO_LAUNCH_RET();
    k_embed_fill<<<gp, 256, 0, stream>>>(a);
    HIMO_LAUNCH_RET();
    dim3 gv(kNumSMs * 4, d->n_frames);
    k_embed_pfn<<<gv, 256, 0, stream>>>(a, g);
    HIMO_LAUNCH_RET();
    k_embed_pfn_big<<<dim3(kNumSMs * 2, d->n_frames), 256, 0, stream>>>(a, g);
    HIMO_LAUNCH_RET();
  } else {
    HIMO_CUDA_RET(cudaMemsetAsync(a.num_voxels, 0, F * sizeof(int), stream));
  }
  return HIMO_OK;
}

// Pointers into the embed workspace, for the stages downstream (decoder gather) and for tests.
extern "C" int himo_embed_views(int n_frames, int n_max, const float* voxel_size, const float* coors_range,
                                void* workspace, himo_embed_view* out) {
  if (!out || !workspace || n_frames <= 0 || n_frames > HIMO_MAX_FRAMES) return HIMO_ERR_ARG;
  const int gx = (int)roundf((coors_range[3] - coors_range[0]) / voxel_size[0]);
  const int gy = (int)roundf((coors_range[4] - coors_range[1]) / voxel_size[1]);
  EmbedArgs a;
  embed_ws_layout(n_frames, n_max, gx * gy / 32 + 1, &a, workspace);
  out->pt4 = (float*)a.pt4; out->bitmap = a.bitmap; out->word_prefix = a.word_prefix;
  out->num_voxels = a.num_voxels; out->rank = a.rank; out->voxel_count = a.count;
  out->seg_start = a.seg_start; out->sorted_idx = a.sorted_idx; out->voxel_feats = a.voxel_feats;
  out->voxel_mean = a.voxel_mean; out->voxel_key = a.voxel_key;
  out->n_words = gx * gy / 32 + 1;
  return HIMO_OK;
}
