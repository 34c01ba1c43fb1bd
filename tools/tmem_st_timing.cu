// Timing probe (not product code): throughput of tcgen05.st 32x32b.x4 / .x16 issued by 4..16 warps at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_st_timing tools/tmem_st_timing.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../synergynet_b200/csrc/tc_common.cuh"
using namespace syn::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

template <int WIDE>
__global__ void __launch_bounds__(512) st_kernel(int reps, long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[16];
  for (int i = 0; i < 16; ++i) r[i] = tid * 16 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int k = 0; k < reps; ++k) {
    if (WIDE) {
      tmem_st16(base + (k & 3) * 16, r);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_st4(base + (k & 3) * 16 + j * 4, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    }
  }
  const long long t1 = clock64();
  tmem_wait_st();
  const long long t2 = clock64();
  if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

int main() {
  long long* d; CK(cudaMalloc(&d, 16));
  for (int wide = 0; wide < 2; ++wide)
    for (int threads : {128, 256, 512})
      for (int reps : {1, 8, 32}) {
        if (wide) st_kernel<1><<<1, threads>>>(reps, d); else st_kernel<0><<<1, threads>>>(reps, d);
        CK(cudaDeviceSynchronize());
        long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        const double bytes = (double)reps * 16 * 4 * threads;
        printf("%s warps=%2d reps=%2d (%.0f KB): issue %lld cycles, +wait %lld cycles -> %.1f B/clk\n", wide ? "x16" : "x4 ", threads / 32, reps,
               bytes / 1024, h[0], h[1], bytes / h[1]);
      }
  return 0;
}
