// Experiment (round 2): tcgen05.mma with MN-major operands read from 64-byte-swizzled tiles that were written as
// K-major tiles of another GEMM -- i.e. can dW = delta^T h consume the row-major [points][features] tiles of delta and h
// directly (M / N = features contiguous, K = points = rows), without the transposed copies FastNSF makes today?
// Tile in shared memory: [chunk c of 32 features][point row][64 bytes], SWIZZLE_64B, as TMA writes a (32 feat x KT pts) box.
// Canonical MN-major layout for SWIZZLE_64B in 16-byte units: ((4,n),(8,k)):((1,LBO),(4,SBO))  (cute mma_traits_sm100.hpp):
//   LBO = byte distance between 32-feature chunks, SBO = byte distance between groups of 8 point rows (= 512).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../himo_b200/csrc/umma.cuh"
using namespace himo;

constexpr int KT = 32;            // points per tile (2 MMAs of K = 16)
constexpr int F = 128;            // features (M and N)
constexpr int CH = F / 32;        // chunks
constexpr int CHB = KT * 64;      // bytes per chunk tile

__device__ __forceinline__ uint64_t desc_mn64(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 4ull << 61;                 // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(128) k_test(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                              float* out, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + CH * CHB;
  uint64_t* bar = (uint64_t*)(sB + CH * CHB); uint64_t* bar2 = bar + 1; uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::mbar_init(bar2, 1); umma::fence_barrier_init(); }
  if (warp == 1) umma::tmem_alloc(tptr, 128);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tm = *tptr;
  if (threadIdx.x == 0) {
    umma::mbar_arrive_expect_tx(bar, 2 * CH * CHB);
    for (int c = 0; c < CH; ++c) {
      umma::tma_load_2d(sA + c * CHB, &tmA, bar, c * 32, 0);
      umma::tma_load_2d(sB + c * CHB, &tmB, bar, c * 32, 0);
    }
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();
    const uint32_t lbo = variant & 1 ? 512u : (uint32_t)CHB, sbo = variant & 1 ? (uint32_t)CHB : 512u;
    // idesc: fp16 x fp16 -> f32, M = N = 128, a_major = b_major = MN (bits 15, 16)
    const uint32_t idesc = umma::idesc_f16kind_f32(128, F, 0, 0) | (1u << 15) | (1u << 16);
    for (int k = 0; k < KT / 16; ++k) {
      const uint64_t ad = desc_mn64(umma::smem_u32(sA) + k * 16 * 64, lbo, sbo);
      const uint64_t bd = desc_mn64(umma::smem_u32(sB) + k * 16 * 64, lbo, sbo);
      umma::mma_bf16_ss(tm, ad, bd, idesc, k ? 1u : 0u);
    }
    umma::mma_commit(bar2);
  }
  umma::mbar_wait(bar2, 0);
  umma::tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < 4; ++c) {
    umma::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * F + c * 32 + j] = __uint_as_float(r[j]);
  }
  umma::tc_fence_before(); __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tm, 128);
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled)p;
  std::vector<__half> A(KT * F), B(KT * F);
  for (int i = 0; i < KT * F; ++i) { A[i] = __float2half((float)((i * 37) % 61 - 30) / 16.f); B[i] = __float2half((float)((i * 53) % 47 - 23) / 16.f); }
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, F * F * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  for (int w = 0; w < 2; ++w) {
    cuuint64_t d[2] = {(cuuint64_t)F, (cuuint64_t)KT}; cuuint64_t s[1] = {(cuuint64_t)F * 2}; cuuint32_t b[2] = {32, (cuuint32_t)KT}; cuuint32_t e[2] = {1, 1};
    if (enc(w ? &tB : &tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w ? (void*)dB : (void*)dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encode failed\n"); return 1; }
  }
  const int smem = 2 * CH * CHB + 64 + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(F * F);
  for (int variant = 0; variant < 2; ++variant) {
    k_test<<<1, 128, smem>>>(tA, tB, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant=%d CUDA error %s\n", variant, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < F; ++m) for (int n = 0; n < F; ++n) {
      double ref = 0; for (int k = 0; k < KT; ++k) ref += (double)__half2float(A[k * F + m]) * (double)__half2float(B[k * F + n]);
      const double er = fabs(ref - O[m * F + n]); if (er > 1e-3) ++bad; maxerr = fmax(maxerr, er);
    }
    printf("MN-major A and B, SWIZZLE_64B, variant=%d (LBO=%d SBO=%d): max_err=%.4g wrong=%d/%d %s\n", variant,
           variant ? 512 : CHB, variant ? CHB : 512, maxerr, bad, F * F, maxerr < 1e-3 ? "OK" : "WRONG");
  }
  return 0;
}
