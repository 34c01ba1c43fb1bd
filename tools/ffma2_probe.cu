// Timing probe (not product code): issue rate of FFMA vs FFMA2 (fma.rn.f32x2), and of tcgen05.ld x8/x16/x32,
// with 4..16 warps on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/ffma2_probe tools/ffma2_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include "../synergynet_b200/csrc/tc_common.cuh"
using namespace syn::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// 16 independent accumulator chains per thread, REPS x 16 FMAs (scalar) or REPS x 16 FFMA2 (packed: 32 FMAs)
template <int PACKED>
__global__ void __launch_bounds__(512) fma_kernel(int reps, float seed, long long* out, float* sink) {
  const int tid = threadIdx.x;
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = seed + tid + i;
  const float w0 = seed * 0.5f, w1 = seed * 0.25f;
  __syncthreads();
  const long long t0 = clock64();
  if (PACKED) {
    unsigned long long p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const unsigned long long w = ((unsigned long long)__float_as_uint(w1) << 32) | __float_as_uint(w0);
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int i = 0; i < 16; ++i) p[i] = ffma2(p[i], w, p[(i + 1) & 15]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[2 * i] = __uint_as_float((unsigned)p[i]); a[2 * i + 1] = __uint_as_float((unsigned)(p[i] >> 32)); }
  } else {
    for (int k = 0; k < reps; ++k) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, a[(i + 1) & 15]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
  sink[tid] = s;
  if (tid == 0) out[0] = t1 - t0;
}

template <int WIDTH>
__global__ void __launch_bounds__(512) ldtm_kernel(int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int k = 0; k < reps; ++k) {
    if (WIDTH == 8) {
      uint32_t r[8], q[8];
      tmem_ld8_async(base + (k & 3) * 16, r);
      tmem_ld8_async(base + (k & 3) * 16 + 8, q);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += __uint_as_float(r[i]) + __uint_as_float(q[i]);
    } else if (WIDTH == 16) {
      float v[16];
      tmem_ld16(base + (k & 3) * 16, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[i];
    } else {
      float v[32];
      tmem_ld32(base + (k & 1) * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += v[i];
    }
  }
  const long long t1 = clock64();
  sink[tid] = acc;
  if (tid == 0) out[0] = t1 - t0;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

int main() {
  long long* d; CK(cudaMalloc(&d, 16));
  float* sink; CK(cudaMalloc(&sink, 4096));
  const int reps = 256;
  for (int packed = 0; packed < 2; ++packed)
    for (int threads : {128, 256, 512}) {
      for (int it = 0; it < 2; ++it) {
        if (packed) fma_kernel<1><<<1, threads>>>(reps, 1.0f, d, sink); else fma_kernel<0><<<1, threads>>>(reps, 1.0f, d, sink);
        CK(cudaDeviceSynchronize());
      }
      long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      const double inst = (double)reps * 16 * (threads / 32);
      printf("%s warps=%2d: %lld cycles, %.3f warp-inst/clk/SM, %.1f fp32 FMA lanes/clk/SM\n", packed ? "FFMA2" : "FFMA ", threads / 32, h,
             inst / h, inst * 32 * (packed ? 2 : 1) / h);
    }
  for (int width : {8, 16, 32})
    for (int threads : {128, 256, 512}) {
      for (int it = 0; it < 2; ++it) {
        if (width == 8) ldtm_kernel<8><<<1, threads>>>(64, d, sink);
        else if (width == 16) ldtm_kernel<16><<<1, threads>>>(64, d, sink);
        else ldtm_kernel<32><<<1, threads>>>(64, d, sink);
        CK(cudaDeviceSynchronize());
      }
      long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      const double bytes = 64.0 * (width == 8 ? 16 : width) * 4 * threads;
      printf("tcgen05.ld x%-2d warps=%2d: %lld cycles -> %.1f B/clk/SM\n", width, threads / 32, h, bytes / h);
    }
  return 0;
}
