// Timing probe (not product code): cost of tcgen05.mma kind::f16 M=128 in SS mode (both operands in shared
// memory, canonical K-major no-swizzle layout) as a function of N and of the number of back-to-back MMAs,
// plus the commit -> mbarrier round trip.  One CTA, one issuing thread, clock64 around issue + wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_timing tools/umma_timing.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../synergynet_b200/csrc/tc_common.cuh"

using namespace syn::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

// Same measurement with the whole warp 0 running the issue loop convergently (descriptors provably
// warp-uniform) and only the MMA itself under elect.sync.
__global__ void __launch_bounds__(128) timing_uniform_kernel(int N, int ksteps, int reps, int batches, long long* out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  for (int i = tid; i < (128 + 256) * 16 * ksteps * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 16 * ksteps * 2;
    const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
    uint32_t phase = 0;
    for (int b = 0; b < batches; ++b) {
      const long long t0 = clock64();
      int ks = 0;
      for (int r = 0; r < reps; ++r) {
        const uint64_t ad = make_smem_desc(a0 + ks * 2 * lboA, lboA, 128), bd = make_smem_desc(b0 + ks * 2 * lboB, lboB, 128);
        if (elect_one()) umma_f16(tmem, ad, bd, idesc, r > 0 ? 1u : 0u);
        ks = (ks + 1 == ksteps) ? 0 : ks + 1;
      }
      const long long t1 = clock64();
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, err);
      phase ^= 1;
      const long long t2 = clock64();
      if (tid == 0) {
        out[2 * b] = t1 - t0;
        out[2 * b + 1] = t2 - t0;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// Mode 2: elect once, then the MMAs back to back from descriptors precomputed in registers (unrolled x6).
__global__ void __launch_bounds__(128) timing_batched_kernel(int N, int ksteps, int reps, int batches, long long* out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  for (int i = tid; i < (128 + 256) * 16 * ksteps * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 16 * 6 * 2;
    const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
    uint64_t ad[6], bd[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      ad[k] = make_smem_desc(a0 + k * 2 * lboA, lboA, 128);
      bd[k] = make_smem_desc(b0 + k * 2 * lboB, lboB, 128);
    }
    uint32_t phase = 0;
    for (int b = 0; b < batches; ++b) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < reps; r += 6) {
#pragma unroll
          for (int k = 0; k < 6; ++k) umma_f16(tmem, ad[k], bd[k], idesc, (r + k) > 0 ? 1u : 0u);
        }
      }
      __syncwarp();
      const long long t1 = clock64();
      if (elect_one()) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), phase, err);
      phase ^= 1;
      const long long t2 = clock64();
      if (tid == 0) {
        out[2 * b] = t1 - t0;
        out[2 * b + 1] = t2 - t0;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// reps MMAs of 128 x N x 16, operands walking over `ksteps` K-slices of a [128 x 16*ksteps] A tile and a
// [N x 16*ksteps] B tile (so consecutive MMAs read different smem, like a real K loop)
__global__ void __launch_bounds__(128) timing_kernel(int N, int ksteps, int reps, int batches, long long* out, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  for (int i = tid; i < (128 + 256) * 16 * ksteps * 2 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 128 * 16 * ksteps * 2;
    const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;      // K-group stride = rows * 16 B
    uint32_t phase = 0;
    for (int b = 0; b < batches; ++b) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const int ks = r % ksteps;
        umma_f16(tmem, make_smem_desc(a0 + ks * 2 * lboA, lboA, 128), make_smem_desc(b0 + ks * 2 * lboB, lboB, 128), idesc,
                 r > 0 ? 1u : 0u);
      }
      const long long t1 = clock64();
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase, err);
      phase ^= 1;
      const long long t2 = clock64();
      out[2 * b] = t1 - t0;
      out[2 * b + 1] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d_out;
  int* d_err;
  CK(cudaMalloc(&d_out, 64 * sizeof(long long)));
  CK(cudaMalloc(&d_err, sizeof(int)));
  CK(cudaMemset(d_err, 0, sizeof(int)));
  CK(cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(timing_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(timing_uniform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int Ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
  const int Rs[] = {0, 6, 18, 72};
  printf("mode N reps ksteps issue_cycles total_cycles cycles_per_mma\n");
  for (int mode = 0; mode < 3; ++mode)
  for (int N : Ns)
    for (int R : Rs) {
      const int ksteps = 6, batches = 8;
      const size_t smem = (size_t)(128 + 256) * 16 * ksteps * 2 + 1024;
      if (mode == 0) timing_kernel<<<1, 128, smem>>>(N, ksteps, R, batches, d_out, d_err);
      else if (mode == 1) timing_uniform_kernel<<<1, 128, smem>>>(N, ksteps, R, batches, d_out, d_err);
      else timing_batched_kernel<<<1, 128, smem>>>(N, ksteps, R, batches, d_out, d_err);
      CK(cudaDeviceSynchronize());
      long long h[64];
      CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
      long long issue = 1LL << 60, total = 1LL << 60;
      for (int b = 2; b < batches; ++b) {           // min over the warm batches
        if (h[2 * b] < issue) issue = h[2 * b];
        if (h[2 * b + 1] < total) total = h[2 * b + 1];
      }
      printf("%s %3d %3d %d %6lld %6lld %.1f\n", mode == 2 ? "batched" : mode ? "warp-uniform" : "one-thread", N, R, ksteps, issue, total, R ? (double)total / R : 0.0);
    }
  int e = 0;
  CK(cudaMemcpy(&e, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  printf("err flag %d\n", e);
  return 0;
}
