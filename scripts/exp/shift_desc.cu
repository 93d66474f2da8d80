// Experiment: can a tcgen05 K-major swizzled A operand start at an arbitrary ROW of a larger TMA-written
// tile (halo reuse across the kx taps of a 3x3 convolution)?  Tests SWIZZLE_64B (64-byte rows) and
// SWIZZLE_128B (128-byte rows) with row shifts 0..3 and the descriptor base_offset field either 0 or
// (start_address >> 7) & 7.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../himo_b200/csrc/umma.cuh"
using namespace himo;

template <int ROWB>
__global__ void __launch_bounds__(128) k_test(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                              float* out, int shift, int bo_mode) {
  constexpr int BK = ROWB / 2, ROWS = 144, N = 64;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + ROWS * ROWB;   // ROWS*ROWB is a multiple of 1024
  uint64_t* bar = (uint64_t*)(sB + N * ROWB); uint64_t* bar2 = bar + 1; uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { umma::mbar_init(bar, 1); umma::mbar_init(bar2, 1); umma::fence_barrier_init(); }
  if (warp == 1) umma::tmem_alloc(tptr, 64);
  umma::tc_fence_before(); __syncthreads(); umma::tc_fence_after();
  const uint32_t tm = *tptr;
  if (threadIdx.x == 0) {
    umma::mbar_arrive_expect_tx(bar, 136 * ROWB + N * ROWB);
    umma::tma_load_2d(sA, &tmA, bar, 0, 0);
    umma::tma_load_2d(sB, &tmB, bar, 0, 0);
    umma::mbar_wait(bar, 0);
    umma::tc_fence_after();
    const uint32_t a_addr = umma::smem_u32(sA) + shift * ROWB;
    uint64_t ad = umma::smem_desc_kmajor<ROWB>(a_addr);
    if (bo_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = umma::smem_desc_kmajor<ROWB>(umma::smem_u32(sB));
    const uint32_t idesc = umma::idesc_f16kind_f32(128, N, 0, 0);
    for (int k = 0; k < BK / 16; ++k) umma::mma_bf16_ss(tm, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
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
  if (warp == 1) umma::tmem_dealloc(tm, 64);
}

template <int ROWB> int run(PFN_cuTensorMapEncodeTiled enc) {
  constexpr int BK = ROWB / 2, N = 64, ROWS = 136;
  std::vector<__half> A(ROWS * BK), B(N * BK);
  for (int i = 0; i < ROWS * BK; ++i) A[i] = __float2half((float)((i * 37) % 61 - 30) / 16.f);
  for (int i = 0; i < N * BK; ++i) B[i] = __float2half((float)((i * 53) % 47 - 23) / 16.f);
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  CUtensorMapSwizzle sw = ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  { cuuint64_t d[2] = {(cuuint64_t)BK, ROWS}; cuuint64_t s[1] = {(cuuint64_t)ROWB}; cuuint32_t b[2] = {(cuuint32_t)BK, ROWS}; cuuint32_t e[2] = {1, 1};
    if (enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1; }
  { cuuint64_t d[2] = {(cuuint64_t)BK, N}; cuuint64_t s[1] = {(cuuint64_t)ROWB}; cuuint32_t b[2] = {(cuuint32_t)BK, N}; cuuint32_t e[2] = {1, 1};
    if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1; }
  const int smem = 144 * ROWB + N * ROWB + 64 + 1024;
  cudaFuncSetAttribute(k_test<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * N);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift = 0; shift < 8; ++shift) {
      k_test<ROWB><<<1, 128, smem>>>(tA, tB, dO, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ROWB=%d bo=%d shift=%d CUDA error %s\n", ROWB, bo, shift, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < BK; ++k) ref += (double)__half2float(A[(m + shift) * BK + k]) * (double)__half2float(B[n * BK + k]);
        maxerr = fmax(maxerr, fabs(ref - O[m * N + n]));
      }
      printf("ROWB=%3d base_offset_mode=%d shift=%d  max_err=%.4g %s\n", ROWB, bo, shift, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled)p;
  return run<64>(enc) | run<128>(enc);
}
