// Bring-up probe (not product code): tcgen05.mma with the A operand in TMEM (written by tcgen05.st as packed
// fp16 pairs) and B in shared memory -- checks the A layout against a CPU product and times back-to-back MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_ts_probe tools/umma_ts_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../synergynet_b200/csrc/tc_common.cuh"

using namespace syn::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

// A [128][K] fp32, B [N][K] fp32 (both exactly representable in fp16), out [128][N]; a_col = TMEM column of A
__global__ void __launch_bounds__(128) ts_kernel(const float* A, const float* Bm, float* out, int N, int K, int a_col,
                                                 int reps, long long* cyc, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  const uint32_t lboB = (uint32_t)N * 16;
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__half*>(smem + (n >> 3) * 128 + (k >> 3) * lboB + (n & 7) * 16 + (k & 7) * 2) = __float2half(Bm[i]);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  // A row `tid` -> TMEM lane tid, columns a_col + k/2
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t r[4];
    for (int j = 0; j < 4; ++j) {
      const __half lo = __float2half(A[tid * K + k0 + 2 * j]), hi = __float2half(A[tid * K + k0 + 2 * j + 1]);
      r[j] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    tmem_st4(tmem + ((uint32_t)(warp * 32) << 16) + a_col + k0 / 2, r[0], r[1], r[2], r[3]);
  }
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const long long t0 = clock64();
    uint64_t bd[6];                                   // K = 96: six K steps, descriptors hoisted out of the loop
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) bd[ks] = make_smem_desc(smem_u32(smem) + ks * 2 * lboB, lboB, 128);
    const uint32_t ta = tmem + a_col;
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) umma_f16_ts(tmem, ta + ks * 8, bd[ks], idesc, (r > 0 || ks > 0) ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, err);
    if (tid == 0) cyc[0] = clock64() - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) out[tid * N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int N : Ns) {
    const int K = 96, a_col = 256;
    std::vector<float> A(128 * K), B(N * K), ref(128 * N), out(128 * N);
    srand(1);
    for (auto& x : A) x = (float)((rand() % 17) - 8) / 8.0f;
    for (auto& x : B) x = (float)((rand() % 17) - 8) / 4.0f;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
        ref[m * N + n] = (float)s;
      }
    float *dA, *dB, *dO;
    long long* dC;
    int* dE;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
    CK(cudaMalloc(&dC, 8)); CK(cudaMalloc(&dE, 4)); CK(cudaMemset(dE, 0, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int reps : {1, 12, 48}) {
      ts_kernel<<<1, 128, N * K * 2 + 1024>>>(dA, dB, dO, N, K, a_col, reps, dC, dE);
      CK(cudaDeviceSynchronize());
      long long cyc;
      int e;
      CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&e, dE, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs(out[i] - reps * ref[i]));
      printf("TS N=%3d K=%d reps=%2d: max|err| %.3g (ref max %.3g) cycles %lld (%.1f / MMA) err_flag %d\n", N, K, reps, maxerr,
             (double)fabs(ref[0]), cyc, (double)cyc / (reps * K / 16), e);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC); cudaFree(dE);
  }
  return 0;
}
