// Experiment: tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form), staged by tcgen05.cp.128x256b from the
// same K-major SWIZZLE_64B shared-memory tile the SS form reads.  Question: does cp + ts-MMA reproduce the SS result
// (so that a_hi can be read from shared memory once for its two products in the split-precision convolution)?
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../himo_b200/csrc/umma.cuh"
using namespace himo;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128) k_test(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                              float* out, int mode) {
  constexpr int ROWB = 64, BK = 32, N = 64;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + 128 * ROWB;
  uint64_t* bar = (uint64_t*)(sB + N * ROWB); uint64_t* bar2 = bar + 1; uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::mbar_init(bar2, 1); umma::fence_barrier_init(); }
  if (warp == 1) umma::tmem_alloc(tptr, 128);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tm = *tptr;
  const uint32_t tA = tm + 64;           // columns [64, 80): two K=16 slices of A, 8 columns each
  if (threadIdx.x == 0) {
    umma::mbar_arrive_expect_tx(bar, 128 * ROWB + N * ROWB);
    umma::tma_load_2d(sA, &tmA, bar, 0, 0);
    umma::tma_load_2d(sB, &tmB, bar, 0, 0);
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();
    const uint64_t ad = umma::smem_desc_kmajor<ROWB>(umma::smem_u32(sA));
    const uint64_t bd = umma::smem_desc_kmajor<ROWB>(umma::smem_u32(sB));
    const uint32_t idesc = umma::idesc_f16kind_f32(128, N, 0, 0);
    for (int k = 0; k < BK / 16; ++k) {
      if (mode == 0) umma::mma_bf16_ss(tm, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
      else {
        tmem_cp_128x256b(tA + 8 * k, ad + (uint64_t)(k * 2));
        mma_f16_ts(tm, tA + 8 * k, bd + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
      }
    }
    umma::mma_commit(bar2);
  }
  umma::mbar_wait(bar2, 0);
  umma::tc_fence_after();
  uint32_t r[32];
  for (int c = 0; c < 2; ++c) {
    umma::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(r[j]);
  }
  umma::tc_fence_before(); __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tm, 128);
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled)p;
  constexpr int ROWB = 64, BK = 32, N = 64, ROWS = 128;
  std::vector<__half> A(ROWS * BK), B(N * BK);
  for (int i = 0; i < ROWS * BK; ++i) A[i] = __float2half((float)((i * 37) % 61 - 30) / 16.f);
  for (int i = 0; i < N * BK; ++i) B[i] = __float2half((float)((i * 53) % 47 - 23) / 16.f);
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  { cuuint64_t d[2] = {(cuuint64_t)BK, ROWS}; cuuint64_t s[1] = {(cuuint64_t)ROWB}; cuuint32_t b[2] = {(cuuint32_t)BK, ROWS}; cuuint32_t e[2] = {1, 1};
    if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1; }
  { cuuint64_t d[2] = {(cuuint64_t)BK, N}; cuuint64_t s[1] = {(cuuint64_t)ROWB}; cuuint32_t b[2] = {(cuuint32_t)BK, N}; cuuint32_t e[2] = {1, 1};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1; }
  const int smem = 128 * ROWB + N * ROWB + 64 + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * N);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dO, 0, 128 * N * 4);
    k_test<<<1, 128, smem>>>(tA, tB, dO, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode=%d CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < BK; ++k) ref += (double)__half2float(A[m * BK + k]) * (double)__half2float(B[n * BK + k]);
      const double err = fabs(ref - O[m * N + n]);
      if (err > 1e-3) ++bad;
      maxerr = fmax(maxerr, err);
    }
    printf("mode=%d (%s)  max_err=%.4g  bad=%d/%d %s\n", mode, mode ? "tcgen05.cp 128x256b + TS mma" : "SS mma", maxerr, bad, 128 * N,
           maxerr < 1e-3 ? "OK" : "WRONG");
    if (maxerr >= 1e-3) { for (int m = 0; m < 4; ++m) { for (int n = 0; n < 6; ++n) printf("%9.3f ", O[m * N + n]); printf("\n"); } }
  }
  return 0;
}
