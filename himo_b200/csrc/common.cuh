// himo_b200/csrc/common.cuh -- shared device/host helpers for libhimo_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define HIMO_OK 0
#define HIMO_ERR_ARG (-1)
#define HIMO_ERR_WORKSPACE (-2)
#define HIMO_ERR_UNSUPPORTED (-3)

// Launch-status helper: the C ABI never throws; >0 return values are cudaError_t codes.
#define HIMO_CUDA_RET(expr)                         \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) return (int)_e;          \
  } while (0)
// Every kernel launch in this library is followed by HIMO_LAUNCH_RET(): it also bumps the
// process-wide launch counter that bench.py reports as "gpu_launches" (diagnostic only).
extern "C" void himo_count_launch_(void);
#define HIMO_LAUNCH_RET()                           \
  do {                                              \
    himo_count_launch_();                           \
    cudaError_t _e = cudaGetLastError();            \
    if (_e != cudaSuccess) return (int)_e;          \
  } while (0)

namespace himo {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Bump allocator over a caller-supplied workspace (C ABI: no hidden allocations).
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  __host__ Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
  template <typename T>
  __host__ T* take(size_t count) {
    off = align_up(off, 256);
    T* r = (T*)(base + off);
    off += count * sizeof(T);
    return r;
  }
  __host__ bool ok() const { return off <= cap && (base != nullptr || off == 0); }
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor in
// the stream is still running; pdl_wait() blocks until that predecessor has completed and its memory is visible
// (a no-op for a normal launch).  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as soon as
// every CTA of this grid has called it (they still block in their own pdl_wait()).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- streaming 128-bit accesses -------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream_i4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ---- warp / block primitives ----------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// Exclusive scan of one int per thread over a block of up to 1024 threads.
// `smem` must hold 33 ints.  Returns the exclusive prefix; *total gets the block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* smem, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  int incl = warp_incl_scan(v);
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? smem[lane] : 0;
    int wi = warp_incl_scan(w);
    smem[lane] = wi - w;
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  int r = smem[warp] + incl - v;
  *total = smem[32];
  __syncthreads();
  return r;
}

// ---- device-wide exclusive scan (int32), decoupled look-back, single pass --------------
// status[t] packs {flag:2 | value:62}: flag 1 = tile aggregate, 2 = inclusive prefix.
// `status` (>= num_tiles u64) and `tile_counter` (1 int) must be zero on entry.
// n is read from *n_dev when n_dev != nullptr (device-sized problems), else n_host.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

template <typename MapFn>
__global__ void __launch_bounds__(kScanThreads)
k_scan_lookback(MapFn map, int* __restrict__ out, int n_host, const int* __restrict__ n_dev,
                unsigned long long* __restrict__ status, int* __restrict__ tile_counter,
                int* __restrict__ total_out) {
  __shared__ int s_scan[33];
  __shared__ int s_tile;
  __shared__ int s_prefix;
  const int n = n_dev ? *n_dev : n_host;
  const int num_tiles = (n + kScanTile - 1) / kScanTile;
  if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1);
  __syncthreads();
  const int tile = s_tile;
  if (tile >= num_tiles) {
    if (tile == 0 && threadIdx.x == 0 && total_out) *total_out = 0;
    return;
  }
  const int base = tile * kScanTile + threadIdx.x * kScanItems;
  int v[kScanItems];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int i = base + k;
    v[k] = i < n ? map(i) : 0;
    sum += v[k];
  }
  int tile_total;
  int excl = block_excl_scan(sum, s_scan, &tile_total);
  if (threadIdx.x == 0) {
    unsigned long long pub = ((tile == 0 ? 2ull : 1ull) << 62) | (unsigned long long)(unsigned)tile_total;
    atomicExch(&status[tile], pub);
    int prefix = 0;
    if (tile > 0) {
      int look = tile - 1;
      while (true) {
        unsigned long long s = atomicAdd(&status[look], 0ull);
        unsigned flag = (unsigned)(s >> 62);
        if (flag == 0) { __nanosleep(20); continue; }
        prefix += (int)(unsigned)(s & 0xffffffffull);
        if (flag == 2) break;
        --look;
      }
      atomicExch(&status[tile], (2ull << 62) | (unsigned long long)(unsigned)(prefix + tile_total));
    }
    s_prefix = prefix;
    if (tile == num_tiles - 1 && total_out) *total_out = prefix + tile_total;
  }
  __syncthreads();
  int run = s_prefix + excl;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
}

struct ScanScratch {
  unsigned long long* status;  // ceil(n_max / kScanTile) entries
  int* tile_counter;           // 1 int (stored right after status)
  static __host__ size_t bytes(long long n_max) {
    return align_up((size_t)(ceil_div_ll(n_max, kScanTile) + 1) * sizeof(unsigned long long) + 16, 256);
  }
};

// Host launcher.  `scratch` must provide ScanScratch::bytes(n_max) bytes.
template <typename MapFn>
inline cudaError_t scan_exclusive(MapFn map, int* out, int n_max, const int* n_dev, int* total_out,
                                  void* scratch, cudaStream_t stream) {
  const long long tiles = ceil_div_ll(n_max > 0 ? n_max : 1, kScanTile);
  size_t sbytes = (size_t)(tiles + 1) * sizeof(unsigned long long) + 16;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sbytes, stream);
  if (e != cudaSuccess) return e;
  unsigned long long* status = (unsigned long long*)scratch;
  int* counter = (int*)(status + tiles + 1);
  k_scan_lookback<<<(unsigned)tiles, kScanThreads, 0, stream>>>(map, out, n_max, n_dev, status, counter,
                                                                total_out);
  himo_count_launch_();
  return cudaGetLastError();
}

struct MapLoadInt {
  const int* p;
  __device__ int operator()(int i) const { return p[i]; }
};
struct MapPopc {
  const unsigned* p;
  __device__ int operator()(int i) const { return __popc(p[i]); }
};

// ---- bitmap-ranked cell index -----------------------------------------------------------
// A dense grid of `n_cells` cells is represented by one bit per cell.  The rank of an
// occupied cell among the occupied cells in ascending key order (= the reference's sorted
// unique_dim order when the key is the row-major (c0,c1,c2) index) is
//     word_prefix[key>>5] + popc(bitmap[key>>5] & ((1<<(key&31))-1)).
// This replaces the reference's full radix sort of the coordinate rows
// (OSF/assets/cuda/mmcv/scatter_points_cuda.cu:24-27) by O(N + n_cells/32) work.
__device__ __forceinline__ int bitmap_rank_lb(const unsigned* __restrict__ bitmap,
                                              const int* __restrict__ word_prefix, long long key) {
  const long long w = key >> 5;
  const unsigned below = (1u << (key & 31)) - 1u;
  return word_prefix[w] + __popc(__ldg(bitmap + w) & below);
}

}  // namespace himo
