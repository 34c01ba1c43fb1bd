// Write-pattern probe for the dense reconstruction output (B,3,N) fp32, N = 53215 (rows are not 16-byte aligned).
// Measures how fast the memory system takes the store stream alone, for the tile shapes the kernel could use:
//   item = F faces x V vertices; a CTA owns F faces (3F output rows) and a band of vertex chunks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dense_store_probe tools/dense_store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void fill_linear(float4* out, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

// policy: 0 st.global, 1 st.global.cs, 2 st.global.wt.  Thread mapping as in the kernel's epilogue: consecutive lanes =
// consecutive vertices, V / 32 warps cover the chunk, the other warps take other faces; the per-thread store loop is
// fully unrolled with constant strides so that the probe is bound by the memory system, not by address arithmetic.
template <int F, int V, int POLICY>
__global__ void __launch_bounds__(512) pattern(float* out, int nver, int batch) {
  constexpr int FG = 512 / V;                  // face groups
  constexpr int FPT = F / FG;                  // faces per thread and item
  const int n_ft = batch / F;
  const int n_vt = (nver + V - 1) / V;
  const int n_bands = max(1, (int)gridDim.x / n_ft);
  const int band_len = (n_vt + n_bands - 1) / n_bands;
  const int ft = blockIdx.x % n_ft, band = blockIdx.x / n_ft;
  if (band >= n_bands) return;
  const int vt_lo = band * band_len, vt_hi = min(vt_lo + band_len, n_vt);
  const int lv = threadIdx.x % V, fg = threadIdx.x / V;
  float* row0 = out + ((size_t)ft * F + fg * FPT) * 3 * nver;
  const size_t n2 = 2 * (size_t)nver;
  for (int vt = vt_lo; vt < vt_hi; ++vt) {
    const int v = vt * V + lv;
    if (v >= nver) continue;
    const float val = (float)vt;
    float* o = row0 + v;
#pragma unroll 16
    for (int f = 0; f < FPT; ++f, o += 3 * (size_t)nver) {
      if (POLICY == 0) { o[0] = val; o[nver] = val; o[n2] = val; }
      else if (POLICY == 1) { __stcs(o, val); __stcs(o + nver, val); __stcs(o + n2, val); }
      else { __stwt(o, val); __stwt(o + nver, val); __stwt(o + n2, val); }
    }
  }
}

// Vector variant: lane -> one 16-byte ALIGNED float4 of a row chunk, whatever the row's own alignment (ld = row stride in
// floats; 53215 puts rows on every 4-byte phase); the <= 3 floats before / after the aligned interior go out as scalars.
template <int F, int V>
__global__ void __launch_bounds__(512) pattern_vec4(float* out, int nver, int ld, int batch) {
  constexpr int TPR = V / 4;                   // threads per row chunk
  constexpr int FG = 512 / TPR;                // rows in flight per sweep
  const int n_ft = batch / F;
  const int n_vt = (nver + V - 1) / V;
  const int n_bands = max(1, (int)gridDim.x / n_ft);
  const int band_len = (n_vt + n_bands - 1) / n_bands;
  const int ft = blockIdx.x % n_ft, band = blockIdx.x / n_ft;
  if (band >= n_bands) return;
  const int vt_lo = band * band_len, vt_hi = min(vt_lo + band_len, n_vt);
  const int j = threadIdx.x % TPR, rg = threadIdx.x / TPR;
  for (int vt = vt_lo; vt < vt_hi; ++vt) {
    const int v0 = vt * V, v1 = min(v0 + V, nver);
    const float val = (float)vt;
#pragma unroll 4
    for (int r = rg; r < 3 * F; r += FG) {
      const size_t row = ((size_t)ft * F * 3 + r) * ld;
      const size_t g0 = row + v0, g1 = row + v1;
      const size_t a0 = (g0 + 3) & ~(size_t)3, a1 = g1 & ~(size_t)3;
      const size_t a = a0 + 4 * (size_t)j;
      if (a + 4 <= a1) *reinterpret_cast<float4*>(out + a) = make_float4(val, val, val, val);
      if (g0 + j < a0) out[g0 + j] = val;                          // head (j < 3)
      if (a1 + j < g1) out[a1 + j] = val;                          // tail (j < 3)
      if (j == TPR - 1 && a + 4 > a1 && a < a1) {}                // (interior always whole float4s)
    }
  }
}

// Window variant: every row chunk is shifted down to the row's own 32-byte sector grid -- [down8(g0), down8(g0) + V) --
// so that ALL stores are whole aligned sectors in the real layout (the kernel would carry the <= 7 floats that fall off the
// end of one chunk into the next chunk of the same row; only the two ends of a CTA's band need scalar stores).
template <int F, int V>
__global__ void __launch_bounds__(512) pattern_win(float* out, int nver, int ld, int batch) {
  constexpr int TPR = V / 4;
  constexpr int FG = 512 / TPR;
  const int n_ft = batch / F;
  const int n_vt = (nver + V - 1) / V;
  const int n_bands = max(1, (int)gridDim.x / n_ft);
  const int band_len = (n_vt + n_bands - 1) / n_bands;
  const int ft = blockIdx.x % n_ft, band = blockIdx.x / n_ft;
  if (band >= n_bands) return;
  const int vt_lo = band * band_len, vt_hi = min(vt_lo + band_len, n_vt);
  const int j = threadIdx.x % TPR, rg = threadIdx.x / TPR;
  for (int vt = vt_lo; vt < vt_hi; ++vt) {
    const int v0 = vt * V, v1 = min(v0 + V, nver);
    const float val = (float)vt;
#pragma unroll 4
    for (int r = rg; r < 3 * F; r += FG) {
      const size_t row = ((size_t)ft * F * 3 + r) * ld;
      const size_t w0 = (row + v0) & ~(size_t)7, w1 = (row + v1) & ~(size_t)7;
      const size_t a = w0 + 4 * (size_t)j;
      if (a + 4 <= w1) *reinterpret_cast<float4*>(out + a) = make_float4(val, val, val, val);
    }
  }
}

template <int F, int V>
void run_win(float* out, int nver, int batch, const auto& timeit) {
  const int n_ft = batch / F;
  for (int grid : {144, 288}) {
    const int g = (grid / n_ft) * n_ft;
    char name[128];
    snprintf(name, sizeof name, "window F=%d V=%d ld=53215 grid=%d", F, V, g);
    timeit(name, [&] { pattern_win<F, V><<<g, 512>>>(out, nver, 53215, batch); });
  }
}

template <int F, int V>
void run_vec(float* out, int nver, int batch, const auto& timeit) {
  const int n_ft = batch / F;
  for (int ld : {53215, 53248}) {
    for (int grid : {144, 288}) {
      const int g = (grid / n_ft) * n_ft;
      if (g == 0) continue;
      char name[128];
      snprintf(name, sizeof name, "vec4 F=%d V=%d ld=%d grid=%d", F, V, ld, g);
      timeit(name, [&] { pattern_vec4<F, V><<<g, 512>>>(out, nver, ld, batch); });
    }
  }
}

template <int F, int V>
void run_shape(float* out, int nver, int batch, const auto& timeit) {
  const int n_ft = batch / F;
  for (int grid : {144, 148, 288, 576}) {
    const int g = (grid / n_ft) * n_ft;
    if (g == 0) continue;
    char name[128];
    snprintf(name, sizeof name, "F=%d V=%d grid=%d st", F, V, g);
    timeit(name, [&] { pattern<F, V, 0><<<g, 512>>>(out, nver, batch); });
    if (grid == 144) {
      snprintf(name, sizeof name, "F=%d V=%d grid=%d st.cs", F, V, g);
      timeit(name, [&] { pattern<F, V, 1><<<g, 512>>>(out, nver, batch); });
      snprintf(name, sizeof name, "F=%d V=%d grid=%d st.wt", F, V, g);
      timeit(name, [&] { pattern<F, V, 2><<<g, 512>>>(out, nver, batch); });
    }
  }
}

int main() {
  const int nver = 53215, batch = 1024;
  const size_t n = (size_t)batch * 3 * nver;
  float* out;
  cudaMalloc(&out, (size_t)batch * 3 * 53248 * 4 + 1024);
  float* flush;
  const size_t flush_n = 256u << 20;
  cudaMalloc(&flush, flush_n);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch) {
    float best = 1e9f, sum = 0;
    const int reps = 10;
    for (int r = 0; r < reps + 2; ++r) {
      cudaMemsetAsync(flush, r, flush_n);
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    printf("%-44s avg %.4f ms  best %.4f ms  %.0f GB/s (avg)\n", name, sum / reps, best, n * 4 / (sum / reps) * 1e-6);
  };
  timeit("linear float4 fill, 148x8 CTAs", [&] { fill_linear<<<148 * 8, 512>>>((float4*)out, n / 4); });
  timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(out, 0, n * 4); });
  run_shape<64, 128>(out, nver, batch, timeit);
  printf("-- scalar stores, rows padded to 53248 floats (every warp store = one aligned line)\n");
  run_shape<64, 128>(out, 53248, batch, timeit);
  printf("-- aligned float4 interior + scalar head / tail\n");
  run_vec<64, 128>(out, nver, batch, timeit);
  run_vec<64, 512>(out, nver, batch, timeit);
  printf("-- whole-sector windows in the real layout\n");
  run_win<64, 128>(out, nver, batch, timeit);
  run_win<64, 256>(out, nver, batch, timeit);
  run_win<64, 512>(out, nver, batch, timeit);
  run_win<32, 128>(out, nver, batch, timeit);
  run_win<128, 128>(out, nver, batch, timeit);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
