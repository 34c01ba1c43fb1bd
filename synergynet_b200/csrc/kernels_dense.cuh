// 3DMM reconstruction on tcgen05 (reference model_building.py:106-139, reconstruct_vertex_62):
//   S[b, 3v+c] = u[3v+c] + sum_k W[3v+c, k] * alpha[b, k]      (k = 40 shape + 10 expression)
//   V[b, i, v] = sum_c P[b, i, c] * S[b, 3v+c] + t[b, i];   V[b, 1, v] = 121 - V[b, 1, v]
// The basis product is a GEMM with M = vertices, N = faces, K = 50 (padded to 64); running it once per
// coordinate plane (x, y, z) puts the three coordinates of vertex v in the SAME TMEM lane, so the
// 3x3 pose transform is per-thread arithmetic and the (B,3,N) output rows are written with fully
// coalesced 128-byte warp stores.  The kernel is HBM-write bound: 638,580 B per face (dense).
//
//   dense_alpha_kernel     params (B,62) -> de-whitened pose (B,12) fp32 + alpha as fp16 hi/lo B tiles
//   dense_recon_tc_kernel  persistent; item = (128-vertex tile, 64-face tile), vertex-tile major; the
//                          96 KB basis tile (3 planes x hi/lo) stays in smem while the CTA walks over
//                          the face tiles
//     warp 16 lane 0: loader + MMA issuer (3 planes x 3 passes x 4 K-steps, N = 64), 2 TMEM buffers
//     warps 0-15:     epilogue in two groups of 8 warps, one per TMEM buffer (lane = vertex; the two
//                     warps of a lane quarter split the 64 faces)
// Split-16x3 precision scheme of kernels_tc.cuh; basis rows are pre-scaled per vertex row and alpha
// per coefficient (both powers of two, folded back exactly in the epilogue / the basis image).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace syn {

constexpr int kDnK = 64;                                  // padded coefficient count
constexpr int kDnFaces = 64;                              // faces per tile (MMA N)
constexpr int kDnAPlane = 128 * kDnK * 2;                 // 16 KB: one plane (hi or lo) of one coordinate
constexpr int kDnATile = 3 * 2 * kDnAPlane;               // 96 KB per 128-vertex tile: [x|y|z][hi|lo]
constexpr int kDnBPlane = kDnFaces * kDnK * 2;            // 8 KB
constexpr int kDnBTile = 2 * kDnBPlane;                   // 16 KB per face tile: [hi|lo]
constexpr int kDnPoseStride = 20;                        // floats per face: [R|t] (12), crop->image affine kx, sx, ky, sy, kz, pad
constexpr int kDnPoseTile = kDnFaces * kDnPoseStride * 4; // 5 KB
constexpr int kDnBSlot = kDnBTile + kDnPoseTile;
constexpr int kDnMetaTile = 128 * 6 * 4;                  // per vertex tile: u[3][128], 1/rowscale[3][128]
constexpr int kDnBSlots = 4;                              // alpha/pose ring: loads run 3 items ahead of the MMAs
constexpr int kDnSmem = kDnATile + 2 * kDnMetaTile + kDnBSlots * kDnBSlot + 1024;
constexpr int kDnEpiWarps = 16;                          // 4 per TMEM lane quarter: 16 faces each
constexpr int kDnThreads = (kDnEpiWarps + 1) * 32;

// ---- pre-pass -----------------------------------------------------------------------------------------
// alpha image: per face tile [hi plane 64 faces x 64 k][lo plane], canonical K-major (SBO 128, LBO 1024);
// pose: (tiles*64, 12) fp32 rows [R|t] (model_building.py:27-29), zero rows past the batch.
__global__ void __launch_bounds__(64) dense_alpha_kernel(const float* __restrict__ params, const float* __restrict__ mean,
                                                         const float* __restrict__ stdv, const float* __restrict__ ascale,
                                                         uint8_t* __restrict__ aimg, float* __restrict__ pose, int batch,
                                                         int whitening, const float* __restrict__ roi5) {
  const int f = threadIdx.x, tile = blockIdx.x;
  const int b = tile * kDnFaces + f;
  float pr[kNumParams];
#pragma unroll
  for (int j = 0; j < kNumParams; ++j) {
    float v = 0.f;
    if (b < batch) {
      v = params[(size_t)b * kNumParams + j];
      if (whitening) v = v * stdv[j] + mean[j];          // model_building.py:117
    }
    pr[j] = v;
  }
  float* prow = pose + (size_t)(tile * kDnFaces + f) * kDnPoseStride;
#pragma unroll
  for (int j = 0; j < 12; ++j) prow[j] = pr[j];
  // crop -> image affine of _predict_vertices (utils/inference.py:127-138): x*kx+sx, y*ky+sy, z*kz (identity if absent)
  const bool has = roi5 != nullptr && b < batch;
  prow[12] = has ? roi5[(size_t)b * 5 + 0] : 1.f; prow[13] = has ? roi5[(size_t)b * 5 + 1] : 0.f;
  prow[14] = has ? roi5[(size_t)b * 5 + 2] : 1.f; prow[15] = has ? roi5[(size_t)b * 5 + 3] : 0.f;
  prow[16] = has ? roi5[(size_t)b * 5 + 4] : 1.f; prow[17] = prow[18] = prow[19] = 0.f;
  uint8_t* hi = aimg + (size_t)tile * kDnBTile + (f >> 3) * 128 + (f & 7) * 16;
#pragma unroll
  for (int kg = 0; kg < kDnK / 8; ++kg) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k0 = kg * 8 + 2 * j, k1 = k0 + 1;
      const float a0 = (k0 < kNumAlpha) ? pr[12 + k0] * ascale[k0] : 0.f;
      const float a1 = (k1 < kNumAlpha) ? pr[12 + k1] * ascale[k1] : 0.f;
      tc::split2_f16(a0, a1, h[j], l[j]);
    }
    *reinterpret_cast<uint4*>(hi + kg * 1024) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(hi + kDnBPlane + kg * 1024) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

struct DenseArgs {
  const uint8_t* basis_img;   // [vertex tiles][x|y|z][hi|lo][128 x 64] canonical (SBO 128, LBO 2048)
  const float* meta;          // [vertex tiles][6][128]: u_x,u_y,u_z, 1/rowscale_x,_y,_z
  const uint8_t* alpha_img;   // from dense_alpha_kernel
  const float* pose;
  float* out;                 // (B,3,nver)
  int batch, nver, n_vtiles, n_ftiles, transform;
  int affine;                 // apply the per-face crop -> image affine stored behind the pose rows
  int stream_stores;          // 1: st.global.cs (evict-first), 0: plain write-back stores (L2 merges neighbouring 512-byte runs)
  long long* trace;           // debug (SYN_DENSE_TRACE): clock64 stamps of CTA 0, 8 events x 64 items x {group 0, group 1, issuer}
  int* err;
};

// Synchronisation (item i uses B slot / TMEM buffer i & 1; `use` = i >> 1 is its per-slot sequence number):
//   bar_a       basis tile + meta landed (one phase per vertex tile)       loader -> issuer, epilogue
//   bar_bfull   alpha + pose tile landed                                    loader -> issuer, epilogue
//   bar_dfull   MMAs of the item complete                                   tcgen05.commit -> epilogue
//   bar_dfree   epilogue done with the item (TMEM buffer, B slot, and -- at a vertex-tile change --
//               the meta rows of its basis tile have been read)              256 arrivals -> issuer
__global__ void __launch_bounds__(kDnThreads, 1) dense_recon_tc_kernel(const DenseArgs p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_a, bar_bfull[kDnBSlots], bar_dfull[2], bar_dfree[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  float* sMeta = reinterpret_cast<float*>(smem + kDnATile);          // 2 slots (vertex-tile load parity)
  uint8_t* sB = smem + kDnATile + 2 * kDnMetaTile;                   // kDnBSlots slots of alpha + pose

  const int tid = threadIdx.x, warp = tid >> 5;
  const int items = p.n_vtiles * p.n_ftiles;
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int it0 = min((int)blockIdx.x * per, items), it1 = min(it0 + per, items);

  if (tid == 0) {
    mbar_init(smem_u32(&bar_a), 1);
    for (int i = 0; i < kDnBSlots; ++i) mbar_init(smem_u32(&bar_bfull[i]), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_dfull[i]), 1);
      mbar_init(smem_u32(&bar_dfree[i]), kDnEpiWarps * 16);     // one group of 8 warps per TMEM buffer
    }
    fence_mbar_init();
  }
  if (warp == kDnEpiWarps) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < kDnEpiWarps) {
    // ------------------------------ epilogue ------------------------------------------------------
    // Two groups of 8 warps: group g owns TMEM buffer g and every item with (i & 1) == g, so that the MMAs
    // of item i+1 (other buffer, other group) and their completion latency overlap this group's work.
    // Inside a group: warp & 3 = TMEM lane quarter, (warp >> 2) & 1 = which 32 of the 64 faces.
    const int grp = warp >> 3, half = (warp >> 2) & 1;
    const int lane_v = tid & 127;
    const int vt0 = it0 / p.n_ftiles;
    int cur_vt = -1;
    float ux = 0.f, uy = 0.f, uz = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
    for (int it = it0 + grp, i = grp; it < it1; it += 2, i += 2) {
      const int vt = it / p.n_ftiles, ft = it - vt * p.n_ftiles;
      const uint32_t use_par = (uint32_t)(i >> 1) & 1;
      const int sb = i % kDnBSlots;
      mbar_wait(smem_u32(&bar_bfull[sb]), (uint32_t)(i / kDnBSlots) & 1, p.err);   // pose tile visible to this thread
      mbar_wait(smem_u32(&bar_dfull[grp]), use_par, p.err);
      tc_fence_after_sync();
      if (vt != cur_vt) {             // new basis tile: its meta rows (landed before the MMAs of this item ran)
        cur_vt = vt;
        mbar_wait(smem_u32(&bar_a), (uint32_t)(vt - vt0) & 1, p.err);
        const float* m = sMeta + ((vt - vt0) & 1) * (kDnMetaTile / 4);
        ux = m[0 * 128 + lane_v]; uy = m[1 * 128 + lane_v]; uz = m[2 * 128 + lane_v];
        ox = m[3 * 128 + lane_v]; oy = m[4 * 128 + lane_v]; oz = m[5 * 128 + lane_v];
      }
      const int v = vt * 128 + lane_v;
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {                            // 2 x 16 faces per thread
        const int fofs = half * 32 + rnd * 16;
        const float* pose = reinterpret_cast<const float*>(sB + sb * kDnBSlot + kDnBTile) + fofs * kDnPoseStride;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + grp * 192 + fofs;
        float sx[16], sy[16], sz[16];
        tmem_ld16x3(trow, trow + 64, trow + 128, sx, sy, sz);   // three loads in flight, one wait
        const int b0 = ft * kDnFaces + fofs;
        if (v < p.nver) {
#pragma unroll
          for (int f = 0; f < 16; ++f) {
            if (b0 + f < p.batch) {
              const float4 r0 = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride);
              const float4 r1 = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride + 4);
              const float4 r2 = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride + 8);
              const float X = fmaf(sx[f], ox, ux), Y = fmaf(sy[f], oy, uy), Z = fmaf(sz[f], oz, uz);
              float vx = fmaf(r0.x, X, fmaf(r0.y, Y, r0.z * Z)) + r0.w;
              float vy = fmaf(r1.x, X, fmaf(r1.y, Y, r1.z * Z)) + r1.w;
              float vz = fmaf(r2.x, X, fmaf(r2.y, Y, r2.z * Z)) + r2.w;
              if (p.transform) vy = (float)(kImg + 1) - vy;          // model_building.py:129,137
              if (p.affine) {                                        // utils/inference.py:131-136, numpy's fp32 mul then add
                const float4 q = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride + 12);
                vx = __fadd_rn(__fmul_rn(vx, q.x), q.y);
                vy = __fadd_rn(__fmul_rn(vy, q.z), q.w);
                vz = __fmul_rn(vz, pose[f * kDnPoseStride + 16]);
              }
              float* o = p.out + (size_t)(b0 + f) * 3 * p.nver + v;
              __stcs(o, vx); __stcs(o + p.nver, vy); __stcs(o + 2 * (size_t)p.nver, vz);   // write-once stream
            }
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&bar_dfree[grp]));
    }
  } else if (warp == kDnEpiWarps) {
    // ------------------------------ loader + MMA issuer -------------------------------------------
    // The whole warp runs this control flow convergently; bulk copies and MMA batches sit under one
    // elect.sync each, so the descriptors come straight from uniform registers (a `tid == X` branch costs
    // ~170 cycles per MMA, tools/umma_timing: 36 MMAs per item made the issuer the bottleneck of round 1).
    const uint32_t idesc = make_idesc_f16(128, kDnFaces);
    const uint32_t d_hi = smem_desc_hi(128);
    const uint32_t a_lo = smem_desc_lo(smem_u32(sA), 2048);
    auto load_b = [&](int it, int s) {
      if (elect_one()) {
        const int ft = it % p.n_ftiles;
        uint8_t* dst = sB + s * kDnBSlot;
        mbar_expect_tx(smem_u32(&bar_bfull[s]), kDnBSlot);
        bulk_g2s(smem_u32(dst), p.alpha_img + (size_t)ft * kDnBTile, kDnBTile, smem_u32(&bar_bfull[s]));
        bulk_g2s(smem_u32(dst + kDnBTile), p.pose + (size_t)ft * kDnFaces * kDnPoseStride, kDnPoseTile, smem_u32(&bar_bfull[s]));
      }
      __syncwarp();
    };
    auto dfree_wait = [&](int j) {                                   // epilogue finished item j (>= 0)
      mbar_wait(smem_u32(&bar_dfree[j & 1]), (uint32_t)(j >> 1) & 1, p.err);
    };
    int cur_vt = -1;
    const int vt0 = it0 / p.n_ftiles;
    for (int k = 0; k < kDnBSlots - 1; ++k)
      if (it0 + k < it1) load_b(it0 + k, k);
    for (int it = it0, i = 0; it < it1; ++it, ++i) {
      const int vt = it / p.n_ftiles;
      const int s = i & 1;
      if (vt != cur_vt) {
        // the basis tile in smem is overwritten: every MMA of the previous tile must be complete and every
        // epilogue thread past its bar_a wait, both implied by the epilogue having finished item i-1
        if (i >= 1) dfree_wait(i - 1);
        cur_vt = vt;
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bar_a), kDnATile + kDnMetaTile);
          bulk_g2s(smem_u32(sA), p.basis_img + (size_t)vt * kDnATile, kDnATile, smem_u32(&bar_a));
          bulk_g2s(smem_u32(sMeta + ((vt - vt0) & 1) * (kDnMetaTile / 4)), p.meta + (size_t)vt * 6 * 128, kDnMetaTile,
                   smem_u32(&bar_a));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar_a), (uint32_t)(vt - vt0) & 1, p.err);
      }
      const int sb = i % kDnBSlots;
      mbar_wait(smem_u32(&bar_bfull[sb]), (uint32_t)(i / kDnBSlots) & 1, p.err);
      if (i >= 2) dfree_wait(i - 2);                               // TMEM buffer s drained
      tc_fence_after_sync();
      const uint32_t b_lo = smem_desc_lo(smem_u32(sB + sb * kDnBSlot), 1024);
      if (elect_one()) {
#pragma unroll
        for (int plane = 0; plane < 3; ++plane) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_off = (plane * 2 + (pass == 2 ? 1 : 0)) * kDnAPlane;   // W: hi,hi,lo
            const uint32_t b_off = (pass == 1 ? kDnBPlane : 0);                     // alpha: hi,lo,hi
#pragma unroll
            for (int ks = 0; ks < kDnK / 16; ++ks)
              umma_f16(tmem + s * 192 + plane * 64, desc64(d_hi, a_lo + ((a_off + ks * 4096) >> 4)),
                       desc64(d_hi, b_lo + ((b_off + ks * 2048) >> 4)), idesc, (pass > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&bar_dfull[s]));
      }
      __syncwarp();
      // prefetch alpha/pose three items ahead into the slot last used by item i-1 (MMA + epilogue done)
      if (it + kDnBSlots - 1 < it1) {
        if (i >= 1) dfree_wait(i - 1);
        load_b(it + kDnBSlots - 1, (i + kDnBSlots - 1) % kDnBSlots);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kDnEpiWarps) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}


// -------------------------------------------------------------------------------------------------------------------
// Face-major variant for the dense mesh (configs[2]).  The kernel above walks vertex-tile major: a CTA writes 512 B
// to each of 64 faces x 3 rows and moves on to OTHER faces, so every DRAM page is touched once per visit (measured:
// 2.1 TB/s of 6.5).  Here a CTA keeps one 64-face tile and walks over consecutive vertex tiles: each of its 192
// output rows grows by 512 contiguous bytes per item, which L2 write-back turns into long DRAM bursts.  The price is
// a new basis tile per item; it streams from L2 (40 MB image, resident) coordinate plane by coordinate plane (32 KB =
// hi + lo of one coordinate) through a 4-slot ring, two planes ahead of the MMAs.
//   bar_pfull[slot]   plane landed                                                    loader -> issuer
//   bar_pempty[slot]  the 12 MMAs that read the plane are complete (tcgen05.commit)  -> loader
//   bar_mfull[slot]   meta rows of item i (slot i % 4) landed                         loader -> epilogue
//   bar_bfull / bar_dfull / bar_dfree as above
constexpr int kFmPlane = 2 * kDnAPlane;                   // 32 KB: [hi|lo] of one coordinate of one vertex tile
#ifndef SYN_FM_SPLIT
#define SYN_FM_SPLIT 1                                    // bulk copies per plane
#endif
// Six plane slots: a 32 KB plane takes ~4 us from request to completion while the output stream saturates the memory
// system (measured timeline, scripts/dense_trace.py), so the bytes in flight pace the kernel.  A CTA works on ONE face
// tile, so its alpha / pose tile is loaded once instead of through a ring -- that is where the two extra slots come from.
constexpr int kFmPSlots = 6, kFmMetaSlots = 4;
constexpr int kFmSmem = kFmPSlots * kFmPlane + kFmMetaSlots * kDnMetaTile + kDnBSlot + 1024;
static_assert(kDnATile == 3 * kFmPlane, "basis tile = three coordinate planes");
static_assert(kFmSmem <= 227 * 1024, "shared memory");

__global__ void __launch_bounds__(kDnThreads, 1) dense_recon_fm_kernel(const DenseArgs p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_pfull[kFmPSlots], bar_pempty[kFmPSlots], bar_mfull[kFmMetaSlots], bar_bfull,
      bar_dfull[2], bar_dfree[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem;                                                                      // plane ring
  float* sMeta = reinterpret_cast<float*>(smem + kFmPSlots * kFmPlane);                   // kFmMetaSlots meta tiles
  uint8_t* sB = smem + kFmPSlots * kFmPlane + kFmMetaSlots * kDnMetaTile;                 // alpha + pose tile of this CTA's face tile

  const int tid = threadIdx.x, warp = tid >> 5;
  // Work split: CTA = (face tile ft, BAND of consecutive vertex tiles); the bands are the same for every face tile and the
  // CTAs of one band are neighbours in the grid, so the n_ftiles CTAs that need a given basis tile ask for it at about the
  // same time and all but the first hit in L2.  (A plain contiguous split of the item list gives every face tile its own
  // band boundaries: ncu showed the 40 MB basis image read 4.4x from DRAM, and the plane loads -- two planes of look-ahead
  // deep -- were what the epilogue warps waited for.)  Consecutive items with the same vertex tile would form a VISIT and
  // share its planes (the machinery below handles runs of any length; walking face tiles in pairs to halve the basis
  // stream was measured slower: the held planes cost a slot of look-ahead), so here every item is its own visit.
  const int n_bands = max(1, (int)gridDim.x / p.n_ftiles);
  const int band_len = (p.n_vtiles + n_bands - 1) / n_bands;
  const int my_ft = (int)blockIdx.x % p.n_ftiles, my_band = (int)blockIdx.x / p.n_ftiles;
  const int vt_lo = min(my_band * band_len, p.n_vtiles), vt_hi = (my_band < n_bands) ? min(vt_lo + band_len, p.n_vtiles) : vt_lo;
  const int it0 = 0, n_items = vt_hi - vt_lo;
  auto decode = [&](int it, int& vt, int& ft) {
    vt = vt_lo + it;
    ft = my_ft;
  };

  if (tid == 0) {
    for (int i = 0; i < kFmPSlots; ++i) { mbar_init(smem_u32(&bar_pfull[i]), 1); mbar_init(smem_u32(&bar_pempty[i]), 1); }
    for (int i = 0; i < kFmMetaSlots; ++i) mbar_init(smem_u32(&bar_mfull[i]), 1);
    mbar_init(smem_u32(&bar_bfull), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_dfull[i]), 1);
      mbar_init(smem_u32(&bar_dfree[i]), kDnEpiWarps * 16);
    }
    fence_mbar_init();
  }
  if (warp == kDnEpiWarps) tmem_alloc<512>(smem_u32(&tmem_base_s));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp < kDnEpiWarps) {
    // ------------------------------ epilogue (two groups of 8 warps, alternate items) -----------------------------
    // both groups walk the whole item list so that they agree on the visit numbering (meta slot = visit % 4)
    const int grp = warp >> 3, half = (warp >> 2) & 1;
    const int lane_v = tid & 127;
    int visit = -1, cur_vt = -1;
    for (int i = 0; i < n_items; ++i) {
      int vt, ft;
      decode(it0 + i, vt, ft);
      if (vt != cur_vt) { cur_vt = vt; ++visit; }
      if ((i & 1) != grp) continue;
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && (tid & 255) == 0 && i < 64;
      if (tr) p.trace[(grp * 64 + i) * 8 + 0] = clock64();
      mbar_wait(smem_u32(&bar_bfull), 0, p.err);                                       // pose tile visible to this thread
      mbar_wait(smem_u32(&bar_mfull[visit % kFmMetaSlots]), (uint32_t)(visit / kFmMetaSlots) & 1, p.err);   // meta rows visible
      if (tr) p.trace[(grp * 64 + i) * 8 + 1] = clock64();
      mbar_wait(smem_u32(&bar_dfull[grp]), (uint32_t)(i >> 1) & 1, p.err);
      tc_fence_after_sync();
      if (tr) p.trace[(grp * 64 + i) * 8 + 2] = clock64();
      const float* m = sMeta + (visit % kFmMetaSlots) * (kDnMetaTile / 4);
      const float ux = m[0 * 128 + lane_v], uy = m[1 * 128 + lane_v], uz = m[2 * 128 + lane_v];
      const float ox = m[3 * 128 + lane_v], oy = m[4 * 128 + lane_v], oz = m[5 * 128 + lane_v];
      const int v = vt * 128 + lane_v;
#pragma unroll
      for (int rnd = 0; rnd < 2; ++rnd) {                            // 2 x 16 faces per thread
        const int fofs = half * 32 + rnd * 16;
        const float* pose = reinterpret_cast<const float*>(sB + kDnBTile) + fofs * kDnPoseStride;
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16) + grp * 192 + fofs;
        float sx[16], sy[16], sz[16];
        tmem_ld16x3(trow, trow + 64, trow + 128, sx, sy, sz);
        if (tr) p.trace[(grp * 64 + i) * 8 + 3 + rnd] = clock64();
        const int b0 = ft * kDnFaces + fofs;
        if (v < p.nver) {
#pragma unroll
          for (int f = 0; f < 16; ++f) {
            if (b0 + f < p.batch) {
              const float4 r0 = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride);
              const float4 r1 = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride + 4);
              const float4 r2 = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride + 8);
              const float X = fmaf(sx[f], ox, ux), Y = fmaf(sy[f], oy, uy), Z = fmaf(sz[f], oz, uz);
              float vx = fmaf(r0.x, X, fmaf(r0.y, Y, r0.z * Z)) + r0.w;
              float vy = fmaf(r1.x, X, fmaf(r1.y, Y, r1.z * Z)) + r1.w;
              float vz = fmaf(r2.x, X, fmaf(r2.y, Y, r2.z * Z)) + r2.w;
              if (p.transform) vy = (float)(kImg + 1) - vy;          // model_building.py:129,137
              if (p.affine) {                                        // utils/inference.py:131-136, numpy's fp32 mul then add
                const float4 q = *reinterpret_cast<const float4*>(pose + f * kDnPoseStride + 12);
                vx = __fadd_rn(__fmul_rn(vx, q.x), q.y);
                vy = __fadd_rn(__fmul_rn(vy, q.z), q.w);
                vz = __fmul_rn(vz, pose[f * kDnPoseStride + 16]);
              }
              float* o = p.out + (size_t)(b0 + f) * 3 * p.nver + v;
              if (p.stream_stores) { __stcs(o, vx); __stcs(o + p.nver, vy); __stcs(o + 2 * (size_t)p.nver, vz); }
              else { o[0] = vx; o[p.nver] = vy; o[2 * (size_t)p.nver] = vz; }
            }
          }
        }
      }
      if (tr) p.trace[(grp * 64 + i) * 8 + 5] = clock64();
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&bar_dfree[grp]));
    }
  } else if (warp == kDnEpiWarps) {
    // ------------------------------ loader + MMA issuer (converged warp, elect.sync) ------------------------------
    const uint32_t idesc = make_idesc_f16(128, kDnFaces);
    const uint32_t d_hi = smem_desc_hi(128);
    // planes are numbered per VISIT (q = 3 * visit + coordinate); the prefetcher scans the item list on its own
    int next_plane = 0;                                               // next plane to request
    int pre_i = 0, pre_vt = -1;                                       // prefetcher: next unscanned item, vertex tile of the visit being requested
    const uint64_t keep = l2_policy_evict_last();                     // the basis image is re-read once per face-tile pair
    auto dfree_wait = [&](int j) { mbar_wait(smem_u32(&bar_dfree[j & 1]), (uint32_t)(j >> 1) & 1, p.err); };
    // Request planes up to (and including) `upto`.  Plane q reuses the slot of plane q - kFmPSlots, whose MMAs must be complete
    // (bar_pempty, consumed strictly in order).  The meta rows of visit v go to slot v % 4, last read by the epilogue of
    // an item at least three items back, which the TMEM hand-over (dfree of item i - 2) has already waited for.
    auto request_planes = [&](int upto) {
      for (; next_plane <= upto; ++next_plane) {
        const int q = next_plane, slot = q % kFmPSlots, v = q / 3, c = q - 3 * v;
        if (c == 0) {                                                // first plane of a new visit: find its vertex tile
          if (pre_i >= n_items) return;                               // no more visits
          int vt, ft;
          decode(it0 + pre_i, vt, ft);
          pre_vt = vt;
          for (++pre_i; pre_i < n_items; ++pre_i) {                   // skip the other items of this visit
            int vt2, ft2;
            decode(it0 + pre_i, vt2, ft2);
            if (vt2 != vt) break;
          }
        }
        if (q >= kFmPSlots) mbar_wait(smem_u32(&bar_pempty[slot]), (uint32_t)(q / kFmPSlots - 1) & 1, p.err);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&bar_pfull[slot]), kFmPlane);
          // SYN_FM_SPLIT bulk copies per plane on the same barrier (measured: a single 32 KB copy takes ~4 us from request
          // to completion while the output stream saturates the memory system, and that latency -- three planes in
          // flight -- is what paces the kernel)
#pragma unroll
          for (int part = 0; part < SYN_FM_SPLIT; ++part)
            bulk_g2s_hint(smem_u32(sP + slot * kFmPlane + part * (kFmPlane / SYN_FM_SPLIT)),
                          p.basis_img + (size_t)pre_vt * kDnATile + (size_t)c * kFmPlane + part * (kFmPlane / SYN_FM_SPLIT),
                          kFmPlane / SYN_FM_SPLIT, smem_u32(&bar_pfull[slot]), keep);
          if (c == 0) {
            const int ms = v % kFmMetaSlots;
            mbar_expect_tx(smem_u32(&bar_mfull[ms]), kDnMetaTile);
            bulk_g2s_hint(smem_u32(sMeta + ms * (kDnMetaTile / 4)), p.meta + (size_t)pre_vt * 6 * 128, kDnMetaTile,
                          smem_u32(&bar_mfull[ms]), keep);
          }
        }
        __syncwarp();
      }
    };
    if (n_items > 0) {                                               // alpha + pose of this CTA's face tile: once
      if (elect_one()) {
        mbar_expect_tx(smem_u32(&bar_bfull), kDnBSlot);
        bulk_g2s(smem_u32(sB), p.alpha_img + (size_t)my_ft * kDnBTile, kDnBTile, smem_u32(&bar_bfull));
        bulk_g2s(smem_u32(sB + kDnBTile), p.pose + (size_t)my_ft * kDnFaces * kDnPoseStride, kDnPoseTile, smem_u32(&bar_bfull));
      }
      __syncwarp();
    }
    request_planes(kFmPSlots - 2);                                   // planes 0..4 in flight before the first MMA
    int visit = -1, cur_vt = -1;
    for (int i = 0; i < n_items; ++i) {
      int vt, ft;
      decode(it0 + i, vt, ft);
      const bool new_visit = vt != cur_vt;
      if (new_visit) { cur_vt = vt; ++visit; }
      bool last_of_visit = true;
      if (i + 1 < n_items) {
        int vt2, ft2;
        decode(it0 + i + 1, vt2, ft2);
        last_of_visit = vt2 != vt;
      }
      const int s = i & 1;
      const bool tr = p.trace != nullptr && blockIdx.x == 0 && (tid & 31) == 0 && i < 64;
      if (tr) p.trace[(128 + i) * 8 + 0] = clock64();
      mbar_wait(smem_u32(&bar_bfull), 0, p.err);
      if (tr) p.trace[(128 + i) * 8 + 1] = clock64();
      if (i >= 2) dfree_wait(i - 2);                                 // TMEM buffer s drained
      if (tr) p.trace[(128 + i) * 8 + 2] = clock64();
      const uint32_t b_lo = smem_desc_lo(smem_u32(sB), 1024);
      for (int c = 0; c < 3; ++c) {
        const int q = 3 * visit + c, slot = q % kFmPSlots;
        // keep up to four planes ahead of the MMAs in flight -- but never ask for a slot whose current plane this visit
        // still needs (a multi-item visit releases its planes with its last item)
        request_planes(last_of_visit ? q + kFmPSlots - 2 : min(q + kFmPSlots - 2, 3 * visit + kFmPSlots - 1));
        if (new_visit) mbar_wait(smem_u32(&bar_pfull[slot]), (uint32_t)(q / kFmPSlots) & 1, p.err);
        if (tr) p.trace[(128 + i) * 8 + 3 + c] = clock64();            // plane c landed (and requests up to q+2 issued)
        tc_fence_after_sync();
        const uint32_t a_lo = smem_desc_lo(smem_u32(sP + slot * kFmPlane), 2048);
        if (elect_one()) {
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_off = (pass == 2 ? kDnAPlane : 0);                     // W: hi,hi,lo
            const uint32_t b_off = (pass == 1 ? kDnBPlane : 0);                     // alpha: hi,lo,hi
#pragma unroll
            for (int ks = 0; ks < kDnK / 16; ++ks)
              umma_f16(tmem + s * 192 + c * 64, desc64(d_hi, a_lo + ((a_off + ks * 4096) >> 4)),
                       desc64(d_hi, b_lo + ((b_off + ks * 2048) >> 4)), idesc, (pass > 0 || ks > 0) ? 1u : 0u);
          }
          if (last_of_visit) umma_commit(smem_u32(&bar_pempty[slot]));   // the plane may be overwritten once these MMAs are done
          if (c == 2) umma_commit(smem_u32(&bar_dfull[s]));
        }
        __syncwarp();
      }
      if (tr) p.trace[(128 + i) * 8 + 6] = clock64();                  // MMAs of the item issued
      if (tr) p.trace[(128 + i) * 8 + 7] = clock64();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kDnEpiWarps) {
    __syncwarp();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace syn
