// Bring-up probe for the tcgen05 building blocks used by the tensor-core engine (not product code):
// checks smem-descriptor semantics (LBO/SBO), the instruction descriptor, TMEM alloc/ld, commit ->
// mbarrier, bulk g2s copies and the bf16x3 split against a CPU reference.  Every wait is bounded.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../synergynet_b200/csrc/tc_common.cuh"

using namespace syn::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct Case { int N, K, variant, split, bulkB; };

// element (r,k) byte offset inside a canonical no-swizzle K-major tile
__host__ __device__ inline uint32_t canon_off(int r, int k, uint32_t lbo, uint32_t sbo) {
  return (r >> 3) * sbo + (k >> 3) * lbo + (r & 7) * 16 + (k & 7) * 2;
}

// A: [128][K] fp32 row-major, Bm: [N][K] fp32 row-major, Bimg: pre-packed canonical image(s) of B
// out: [128][N] fp32.  split=0: single bf16 pass; split=1: bf16x3 (hi*hi + hi*lo + lo*hi)
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* Bm, const uint8_t* Bimg, float* out,
                                                    int N, int K, uint32_t lboA, uint32_t sboA, uint32_t lboB,
                                                    uint32_t sboB, int dcol, int split, int bulkB, int* err) {
  const int swap_fields = 0;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t a_bytes = 128 * K * 2, b_bytes = N * K * 2;
  uint8_t* sAh = smem;
  uint8_t* sAl = sAh + a_bytes;
  uint8_t* sBh = sAl + a_bytes;
  uint8_t* sBl = sBh + b_bytes;

  if (tid == 0) {
    mbar_init(smem_u32(&bar_mma), 1);
    mbar_init(smem_u32(&bar_tma), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_base_s));
  // fill A (thread = row)
  for (int k = 0; k < K; ++k) {
    __nv_bfloat16 hi, lo;
    split_bf16(A[tid * K + k], hi, lo);
    *reinterpret_cast<__nv_bfloat16*>(sAh + canon_off(tid, k, lboA, sboA)) = hi;
    *reinterpret_cast<__nv_bfloat16*>(sAl + canon_off(tid, k, lboA, sboA)) = lo;
  }
  if (!bulkB) {
    for (int i = tid; i < N * K; i += 128) {
      const int n = i / K, k = i % K;
      __nv_bfloat16 hi, lo;
      split_bf16(Bm[i], hi, lo);
      *reinterpret_cast<__nv_bfloat16*>(sBh + canon_off(n, k, lboB, sboB)) = hi;
      *reinterpret_cast<__nv_bfloat16*>(sBl + canon_off(n, k, lboB, sboB)) = lo;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s + dcol;

  if (bulkB) {
    if (tid == 0) {
      mbar_expect_tx(smem_u32(&bar_tma), 2 * b_bytes);
      bulk_g2s(smem_u32(sBh), Bimg, b_bytes, smem_u32(&bar_tma));
      bulk_g2s(smem_u32(sBl), Bimg + b_bytes, b_bytes, smem_u32(&bar_tma));
    }
    mbar_wait(smem_u32(&bar_tma), 0, err);
  }

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const int passes = split ? 3 : 1;
    uint32_t acc = 0;
    for (int p = 0; p < passes; ++p) {
      const uint8_t* a = (p == 2) ? sAl : sAh;   // hi*hi, hi*lo, lo*hi
      const uint8_t* b = (p == 1) ? sBl : sBh;
      for (int k0 = 0; k0 < K; k0 += 16) {
        const uint32_t aaddr = smem_u32(a) + (k0 >> 3) * lboA;
        const uint32_t baddr = smem_u32(b) + (k0 >> 3) * lboB;
        const uint64_t ad = swap_fields ? make_smem_desc(aaddr, sboA, lboA) : make_smem_desc(aaddr, lboA, sboA);
        const uint64_t bd = swap_fields ? make_smem_desc(baddr, sboB, lboB) : make_smem_desc(baddr, lboB, sboB);
        umma_f16(tmem, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    umma_commit(smem_u32(&bar_mma));
  }
  mbar_wait(smem_u32(&bar_mma), 0, err);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) out[tid * N + c0 + j] = v[j];
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem - dcol);
}

static float bf16_round(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  r &= 0xFFFF0000u;
  float y;
  memcpy(&y, &r, 4);
  return y;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d\n", prop.name, prop.major, prop.minor);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int NK[][2] = {{16, 16}, {48, 32}, {64, 64}, {96, 32}, {32, 144}, {256, 96}, {160, 128}, {240, 64}};
  int n_bad = 0;
  for (auto& nk : NK) {
    const int N = nk[0], K = nk[1];
    std::vector<float> A(128 * K), B(N * K), ref(128 * N), ref64(128 * N);
    srand(N * 131 + K);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        double d = 0.0;
        for (int k = 0; k < K; ++k) {
          s += bf16_round(A[m * K + k]) * bf16_round(B[n * K + k]);
          d += (double)A[m * K + k] * (double)B[n * K + k];
        }
        ref[m * N + n] = s;
        ref64[m * N + n] = (float)d;
      }
    float *dA, *dB, *dO;
    uint8_t* dImg;
    int* dErr;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, ref.size() * 4));
    CK(cudaMalloc(&dImg, (size_t)N * K * 4)); CK(cudaMalloc(&dErr, 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int variant = 0; variant < 4; ++variant)
      for (int split = 0; split < 2; ++split)
        for (int bulk = 0; bulk < 2; ++bulk) {
          const int lay = variant & 1, swp = 0;
          const int dcol = (variant >> 1) ? (N <= 160 ? 48 + 16 * (N % 3) : 16) : 0;
          if (dcol && (!split || !bulk)) continue;
          uint32_t lboA, sboA, lboB, sboB;
          if (lay == 0) { lboA = 128; sboA = (K / 8) * 128; lboB = 128; sboB = (K / 8) * 128; }
          else { sboA = 128; lboA = (128 / 8) * 128; sboB = 128; lboB = (N / 8) * 128; }
          // packed image of B (hi plane then lo plane)
          std::vector<uint16_t> img((size_t)N * K * 2, 0);
          for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
              const float x = B[n * K + k], hi = bf16_round(x), lo = bf16_round(x - hi);
              uint32_t uh, ul;
              memcpy(&uh, &hi, 4); memcpy(&ul, &lo, 4);
              const uint32_t off = canon_off(n, k, lboB, sboB) / 2;
              img[off] = (uint16_t)(uh >> 16);
              img[(size_t)N * K + off] = (uint16_t)(ul >> 16);
            }
          CK(cudaMemcpy(dImg, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
          CK(cudaMemset(dErr, 0, 4));
          CK(cudaMemset(dO, 0xFF, ref.size() * 4));
          const size_t smem = 2 * (128 * K * 2) + 2 * (N * K * 2) + 1024;
          probe_kernel<<<1, 128, smem>>>(dA, dB, dImg, dO, N, K, lboA, sboA, lboB, sboB, dcol, split, bulk, dErr);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("N=%d K=%d v=%d: kernel error %s\n", N, K, variant, cudaGetErrorString(e)); return 3; }
          std::vector<float> out(ref.size());
          int herr = 0;
          CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
          CK(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
          double maxerr = 0, maxref = 0;
          const std::vector<float>& r = split ? ref64 : ref;
          for (size_t i = 0; i < out.size(); ++i) {
            double d = fabs((double)out[i] - (double)r[i]);
            if (!(d == d)) d = 1e30;
            if (d > maxerr) maxerr = d;
            if (fabs(r[i]) > maxref) maxref = fabs(r[i]);
          }
          const double rel = maxerr / maxref;
          const bool expect_ok = (swp == 0);
          const bool ok = rel < (split ? 3e-5 : 2e-5);
          printf("N=%3d K=%3d layout=%d dcol=%d split=%d bulkB=%d : rel err %.3e %s%s\n", N, K, lay, dcol, split,
                 bulk, rel, ok ? "OK" : "BAD", herr ? " [WAIT TIMEOUT]" : "");
          if (expect_ok && !ok) ++n_bad;
        }
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dImg); cudaFree(dErr);
  }
  printf("probe summary: %d unexpected failures\n", n_bad);
  return n_bad ? 1 : 0;
}
