// Sim3DR on the GPU: vertex normals, per-vertex lighting and z-buffer rasterisation of a
// BATCH of meshes that share one triangle list -- the consumer of the dense vertices `syn_reconstruct_image` leaves in
// HBM, which it reads in place (plane-major (B,3,N) or the reference's interleaved (N,3): element strides).
//
// The reference (Sim3DR/lib/rasterize_kernel.cpp) is one serial loop over triangles with a read-modify-write depth
// buffer.  Its result per pixel is "the triangle of greatest interpolated depth, first one on ties" (strict `>` test,
// :241), which is order-free once written as the maximum of a 64-bit key (render_math.h):
//   pass 1  raster_depth_kernel    one thread per (mesh, triangle); atomicMax of the key over the pixels of its bounding
//                                  box; bounding boxes above 64 pixels are walked by the whole warp
//   pass 2  raster_resolve_kernel  one thread per pixel: the LAST mesh that covers it wins (the reference draws meshes
//                                  one after the other onto the same image, utils/render.py:41-45, each with a fresh
//                                  depth buffer), barycentric weights recomputed from the winning triangle, colours
//                                  interpolated and written as the reference's (unsigned char) expression.
// Vertex normals are the sum of the incident face normals IN TRIANGLE ORDER (:189-199): a per-vertex incidence list,
// ascending by construction (syn_mesh_incidence_host), replaces the scatter loop, so the float sums associate exactly
// as the reference's do.  All arithmetic comes from render_math.h (no FMA contraction): normals and rasterisation are
// bit-exact against the reference; lighting is numpy float32 arithmetic except powf (see rmath::powi).
#pragma once
#include "common.cuh"
#include "render_math.h"

namespace syn {

struct MeshView {        // B meshes over one topology
  const float* v;        // coordinate k of vertex i of mesh b: v[b * sb + i * sv + k * sc]
  long long sb;
  int sv, sc;
  int nver, batch;
};

__device__ __forceinline__ void load_vertex(const MeshView& m, int b, int i, float* p) {
  const float* q = m.v + (size_t)b * m.sb + (size_t)i * m.sv;
  p[0] = __ldg(q);
  p[1] = __ldg(q + m.sc);
  p[2] = __ldg(q + 2 * m.sc);
}

// ---- normals -------------------------------------------------------------------------------------------------------
__global__ void tri_normal_kernel(MeshView m, const int32_t* __restrict__ tri, int ntri, float* __restrict__ tn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= ntri) return;
  const int i0 = __ldg(tri + 3 * i), i1 = __ldg(tri + 3 * i + 1), i2 = __ldg(tri + 3 * i + 2);
  float n[3] = {0.f, 0.f, 0.f};
  if ((unsigned)i0 < (unsigned)m.nver && (unsigned)i1 < (unsigned)m.nver && (unsigned)i2 < (unsigned)m.nver) {
    float p0[3], p1[3], p2[3];
    load_vertex(m, b, i0, p0);
    load_vertex(m, b, i1, p1);
    load_vertex(m, b, i2, p2);
    rmath::tri_normal(p0, p1, p2, n);
  }
  float* o = tn + ((size_t)b * ntri + i) * 3;
  o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
}

// inc_start (nver + 1), inc_tri (3 * ntri): triangles incident to each vertex, ascending, one entry per corner
__global__ void vertex_normal_kernel(int nver, int ntri, const float* __restrict__ tn, const int32_t* __restrict__ inc_start,
                                     const int32_t* __restrict__ inc_tri, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= nver) return;
  float n[3] = {0.f, 0.f, 0.f};
  const int e0 = __ldg(inc_start + i), e1 = __ldg(inc_start + i + 1);
  const float* base = tn + (size_t)b * ntri * 3;
  for (int e = e0; e < e1; ++e) {
    const float* t = base + (size_t)__ldg(inc_tri + e) * 3;
    n[0] = rmath::add(n[0], t[0]);
    n[1] = rmath::add(n[1], t[1]);
    n[2] = rmath::add(n[2], t[2]);
  }
  rmath::normalize3(n);
  float* o = out + ((size_t)b * nver + i) * 3;
  o[0] = n[0]; o[1] = n[1]; o[2] = n[2];
}

// ---- lighting ------------------------------------------------------------------------------------------------------
// per-mesh coordinate extremes as order-preserving unsigneds; stats[b][0..2] = max of ~ordered(x) (i.e. the minimum),
// stats[b][3..5] = max of ordered(x); zero-initialised by the caller (cudaMemsetAsync)
__global__ void mesh_extent_kernel(MeshView m, unsigned* __restrict__ stats) {
  const int b = blockIdx.y;
  unsigned lo[3] = {0u, 0u, 0u}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.nver; i += gridDim.x * blockDim.x) {
    float p[3];
    load_vertex(m, b, i, p);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned o = rmath::float_ordered(p[k]);
      lo[k] = max(lo[k], ~o);
      hi[k] = max(hi[k], o);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    lo[k] = __reduce_max_sync(0xFFFFFFFFu, lo[k]);
    hi[k] = __reduce_max_sync(0xFFFFFFFFu, hi[k]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicMax(stats + b * 6 + k, lo[k]);
      atomicMax(stats + b * 6 + 3 + k, hi[k]);
    }
  }
}

// light[b][i][0..2] (Sim3DR/lighting.py:37-66); with a texture (nver,3): colours = texture * light (:74)
__global__ void vertex_light_kernel(MeshView m, const float* __restrict__ normals, const unsigned* __restrict__ stats,
                                    rmath::LightCfg cfg, const float* __restrict__ texture, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= m.nver) return;
  rmath::NormStats s;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s.mn[k] = rmath::ordered_float(~stats[b * 6 + k]);
    s.mx[k] = rmath::ordered_float(stats[b * 6 + 3 + k]);
  }
  float p[3], l[3];
  load_vertex(m, b, i, p);
  const float* n = normals + ((size_t)b * m.nver + i) * 3;
  const float nn[3] = {n[0], n[1], n[2]};
  rmath::vertex_light(p, nn, s, cfg, l);
  float* o = out + ((size_t)b * m.nver + i) * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) o[k] = texture ? rmath::mul(__ldg(texture + (size_t)i * 3 + k), l[k]) : l[k];
}

// ---- rasterisation -------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool load_tri(const MeshView& m, int b, const int32_t* __restrict__ tri, int i, int w, int h,
                                         rmath::TriSetup& t) {
  const int i0 = __ldg(tri + 3 * i), i1 = __ldg(tri + 3 * i + 1), i2 = __ldg(tri + 3 * i + 2);
  if ((unsigned)i0 >= (unsigned)m.nver || (unsigned)i1 >= (unsigned)m.nver || (unsigned)i2 >= (unsigned)m.nver) return false;
  float p[3];
  load_vertex(m, b, i0, p); t.x0 = p[0]; t.y0 = p[1]; t.z0 = p[2];
  load_vertex(m, b, i1, p); t.x1 = p[0]; t.y1 = p[1]; t.z1 = p[2];
  load_vertex(m, b, i2, p); t.x2 = p[0]; t.y2 = p[1]; t.z2 = p[2];
  return rmath::tri_setup(t, w, h);
}

constexpr int kRasterSmallBox = 64;     // bounding boxes up to this many pixels are walked by the owning thread

// keys: (B, h, w) uint64, zero = empty.  One thread per (mesh, triangle).
__global__ void raster_depth_kernel(MeshView m, const int32_t* __restrict__ tri, int ntri, int w, int h,
                                    unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  rmath::TriSetup t;
  const bool live = (i < ntri) && load_tri(m, b, tri, i, w, h, t);
  unsigned long long* kb = keys + (size_t)b * w * h;
  int bw = 0, area = 0;
  if (live) { bw = t.xmax - t.xmin + 1; area = bw * (t.ymax - t.ymin + 1); }
  if (live && area <= kRasterSmallBox) {
    for (int y = t.ymin; y <= t.ymax; ++y)
      for (int x = t.xmin; x <= t.xmax; ++x) {
        uint64_t key;
        if (rmath::pixel_key(t, (uint32_t)i, x, y, key)) atomicMax(kb + (size_t)y * w + x, (unsigned long long)key);
      }
  }
  // large boxes: the warp walks them together, one triangle at a time
  unsigned big = __ballot_sync(0xFFFFFFFFu, live && area > kRasterSmallBox);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    rmath::TriSetup s;
    s.x0 = __shfl_sync(0xFFFFFFFFu, t.x0, src); s.y0 = __shfl_sync(0xFFFFFFFFu, t.y0, src); s.z0 = __shfl_sync(0xFFFFFFFFu, t.z0, src);
    s.x1 = __shfl_sync(0xFFFFFFFFu, t.x1, src); s.y1 = __shfl_sync(0xFFFFFFFFu, t.y1, src); s.z1 = __shfl_sync(0xFFFFFFFFu, t.z1, src);
    s.x2 = __shfl_sync(0xFFFFFFFFu, t.x2, src); s.y2 = __shfl_sync(0xFFFFFFFFu, t.y2, src); s.z2 = __shfl_sync(0xFFFFFFFFu, t.z2, src);
    s.xmin = __shfl_sync(0xFFFFFFFFu, t.xmin, src); s.ymin = __shfl_sync(0xFFFFFFFFu, t.ymin, src);
    const int sbw = __shfl_sync(0xFFFFFFFFu, bw, src), sarea = __shfl_sync(0xFFFFFFFFu, area, src);
    const int si = __shfl_sync(0xFFFFFFFFu, i, src);
    for (int q = lane; q < sarea; q += 32) {
      const int y = s.ymin + q / sbw, x = s.xmin + q % sbw;
      uint64_t key;
      if (rmath::pixel_key(s, (uint32_t)si, x, y, key)) atomicMax(kb + (size_t)y * w + x, (unsigned long long)key);
    }
  }
}

// One thread per pixel.  image (h, w, c) uint8 is updated in place; depth_out (optional, (B,h,w) fp32) receives the
// reference's depth buffer of every mesh (-1e8 where nothing was drawn).
__global__ void raster_resolve_kernel(MeshView m, const int32_t* __restrict__ tri, const float* __restrict__ colors, int c,
                                      int w, int h, float alpha, int reverse, const unsigned long long* __restrict__ keys,
                                      unsigned char* __restrict__ image, float* __restrict__ depth_out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const size_t pix = (size_t)y * w + x;
  bool drawn = false;
  for (int b = m.batch - 1; b >= 0; --b) {
    const unsigned long long key = keys[(size_t)b * w * h + pix];
    if (depth_out) depth_out[(size_t)b * w * h + pix] = key ? rmath::key_depth(key) : rmath::kDepthInit;
    if (key == 0ull || drawn) continue;
    drawn = true;                       // later meshes overwrite earlier ones (alpha == 1)
    const int i = (int)rmath::key_tri(key);
    const int i0 = __ldg(tri + 3 * i), i1 = __ldg(tri + 3 * i + 1), i2 = __ldg(tri + 3 * i + 2);
    float p0[3], p1[3], p2[3];
    load_vertex(m, b, i0, p0);
    load_vertex(m, b, i1, p1);
    load_vertex(m, b, i2, p2);
    const rmath::Bary bw = rmath::barycentric((float)x, (float)y, p0[0], p0[1], p1[0], p1[1], p2[0], p2[1]);
    const float* cb = colors + (size_t)b * m.nver * c;
    unsigned char* dst = image + ((size_t)(reverse ? h - 1 - y : y) * w + x) * c;
    for (int k = 0; k < c; ++k) {
      const float pc = rmath::interp(bw, __ldg(cb + (size_t)i0 * c + k), __ldg(cb + (size_t)i1 * c + k), __ldg(cb + (size_t)i2 * c + k));
      dst[k] = rmath::blend_u8(dst[k], alpha, pc);
    }
    if (!depth_out) break;
  }
}

// ---- NMS (FaceBoxes/utils/nms/cpu_nms.pyx:17-68, py_cpu_nms.py:10-38) -----------------------
// dets (n, 5) fp32 [x1 y1 x2 y2 score], ALREADY in the order the greedy loop visits them (descending score).
// mask[i][j / 64] bit (j % 64) = box j (> i) is suppressed by box i.  ge != 0: `ovr >= thresh` in double (cpu_nms.pyx:65,
// thresh is a C double there); ge == 0: py_cpu_nms keeps `ovr <= thresh` in float32, i.e. suppresses on `ovr > thresh`.
__global__ void nms_mask_kernel(const float* __restrict__ dets, int n, double thresh, int ge, unsigned long long* __restrict__ mask) {
  const int words = (n + 63) / 64;
  const int i = blockIdx.y * blockDim.y + threadIdx.y, wj = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || wj >= words) return;
  unsigned long long bits = 0ull;
  const int j0 = wj * 64;
  if (j0 + 63 > i) {
    const float a[4] = {dets[i * 5], dets[i * 5 + 1], dets[i * 5 + 2], dets[i * 5 + 3]};
    const float area_a = rmath::box_area(a[0], a[1], a[2], a[3]);
    const float thr_f = (float)thresh;
    for (int jj = 0; jj < 64; ++jj) {
      const int j = j0 + jj;
      if (j <= i || j >= n) continue;
      const float bq[4] = {dets[j * 5], dets[j * 5 + 1], dets[j * 5 + 2], dets[j * 5 + 3]};
      const float ovr = rmath::box_overlap(a, area_a, bq, rmath::box_area(bq[0], bq[1], bq[2], bq[3]));
      const bool sup = ge ? ((double)ovr >= thresh) : (ovr > thr_f);
      if (sup) bits |= 1ull << jj;
    }
  }
  mask[(size_t)i * words + wj] = bits;
}

// The greedy scan itself, 64 boxes (one mask word) at a time, one CTA:
//   A  warp 0 resolves the block: the 64 x 64 diagonal of the bit matrix sits in two registers per lane, box b of the
//      block survives iff its bit in `removed` is still clear when its turn comes, and then its diagonal word joins the
//      running word -- 64 shuffle steps, no memory traffic;
//   B  all threads OR the rows of the boxes that survived into `removed` for the words to the right of the block, four
//      independent loads in flight per thread (a box-by-box scan would pay one dependent L2 round trip per kept box).
// keep (n) int32 receives the kept indices in visiting order, *n_keep their number: the serial greedy list, exactly.
constexpr int kNmsScanThreads = 1024;
__global__ void __launch_bounds__(kNmsScanThreads) nms_scan_kernel(const unsigned long long* __restrict__ mask, int n,
                                                                   int32_t* __restrict__ keep, int32_t* __restrict__ n_keep) {
  extern __shared__ unsigned long long removed[];          // words entries
  __shared__ int rows[64];
  __shared__ int n_rows, total;
  const int words = (n + 63) / 64, tid = threadIdx.x, lane = tid & 31;
  for (int wq = tid; wq < words; wq += kNmsScanThreads) removed[wq] = 0ull;
  if (tid == 0) total = 0;
  __syncthreads();
  for (int blk = 0; blk < words; ++blk) {
    const int base = blk * 64, cnt = min(64, n - base);
    if (tid < 32) {
      const unsigned long long d0 = (lane < cnt) ? mask[(size_t)(base + lane) * words + blk] : 0ull;
      const unsigned long long d1 = (lane + 32 < cnt) ? mask[(size_t)(base + 32 + lane) * words + blk] : 0ull;
      unsigned long long cur = removed[blk], keepbits = 0ull;
      for (int b = 0; b < cnt; ++b) {
        const unsigned long long db = __shfl_sync(0xFFFFFFFFu, (b < 32) ? d0 : d1, b & 31);
        if (!((cur >> b) & 1ull)) { keepbits |= 1ull << b; cur |= db; }
      }
      const int t0 = total;
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const int b = lane + 32 * hlf;
        if ((keepbits >> b) & 1ull) {
          const int pos = __popcll(keepbits & ((1ull << b) - 1ull));
          keep[t0 + pos] = base + b;
          rows[pos] = base + b;
        }
      }
      __syncwarp();
      if (lane == 0) { n_rows = __popcll(keepbits); total = t0 + __popcll(keepbits); }
    }
    __syncthreads();
    const int nr = n_rows, rem = words - blk - 1;
    const int pairs = nr * rem;
    for (int p0 = tid; p0 < pairs; p0 += 4 * kNmsScanThreads) {
      unsigned long long v[4];
      int wq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kNmsScanThreads;
        v[u] = 0ull; wq[u] = 0;
        if (p < pairs) {
          const int r = p / rem;
          wq[u] = blk + 1 + (p - r * rem);
          v[u] = mask[(size_t)rows[r] * words + wq[u]];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (v[u]) atomicOr(&removed[wq[u]], v[u]);
    }
    __syncthreads();
  }
  if (tid == 0) *n_keep = total;
}

}  // namespace syn
