// Which (lane, column) does each register of tcgen05.ld.16x256b hold?  Four warps fill 32 columns of TMEM with
// lane * 1000 + column through the 32x32b shape (thread = lane), then warp 1 reads its lane quarter back with 16x256b.x2
// at lane offsets 0 and 16 and prints what every thread received.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Isynergynet_b200/csrc -o tools/tmem_ld_shapes.bin tools/tmem_ld_shapes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace syn::tc;

__global__ void __launch_bounds__(128) probe(int* out) {
  __shared__ uint32_t tbase_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<32>(smem_u32(&tbase_s));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tbase_s;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const int L = warp * 32 + lane;
  for (int c = 0; c < 32; c += 4)
    tmem_st4(trow + c, L * 1000 + c, L * 1000 + c + 1, L * 1000 + c + 2, L * 1000 + c + 3);
  tmem_wait_st();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (warp == 1) {
    for (int half = 0; half < 2; ++half) {
      uint32_t r[8];
      const uint32_t ta = tmem + ((uint32_t)(warp * 32 + half * 16) << 16) + 8;     // columns 8..23
      asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(ta)
                   : "memory");
      tmem_wait_ld();
      for (int j = 0; j < 8; ++j) out[(half * 32 + lane) * 8 + j] = (int)r[j];
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

int main() {
  int* d;
  cudaMalloc(&d, 64 * 8 * sizeof(int));
  cudaMemset(d, 0xff, 64 * 8 * sizeof(int));
  probe<<<1, 128>>>(d);
  int h[64 * 8];
  cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("status: %s (warp 1 = lanes 32..63, columns 8..23 requested; value = lane * 1000 + column)\n", cudaGetErrorString(e));
  for (int half = 0; half < 2; ++half)
    for (int t = 0; t < 32; ++t) {
      printf("half %d thread %2d:", half, t);
      for (int j = 0; j < 8; ++j) printf(" %6d", h[(half * 32 + t) * 8 + j]);
      printf("\n");
    }
  return 0;
}
