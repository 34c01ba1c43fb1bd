// Fused inverted-residual block on tcgen05:  1x1 expand (+BN+ReLU6) -> 3x3 depthwise (+BN+ReLU6)
// -> 1x1 project (+BN)(+skip), reference backbone_nets/mobilenetv2_backbone.py:45-74, in ONE
// kernel, so the 6x-expanded hidden tensor (up to 1.38 MB per face) never touches HBM.
// The same kernel, with an im2col loader, fuses the 3x3/s2 stem conv with block 1
// (mobilenetv2_backbone.py:127 + features[1]).
//
// Work item (tile) = RO output rows of one face (large maps) or FACES whole faces (8x8 / 4x4 maps).
// Per tile the block input is converted once to fp16 hi/lo and stored in TMEM (XA, tcgen05.st); then, for
// each chunk of NC hidden channels:
//   GEMM1  D1[t] (128 x NC, TMEM)  = XA[t] (TMEM, TS mode) * W1c^T (smem)                      (tensor)
//   EPI1   Hs[pixel][NC] (fp32, smem, zero halo) = relu6(s1 * D1 + b1) / 6  (one FFMA.SAT)   (CUDA)
//   DW     A2[out pixel][NC] (fp16 hi/lo, smem)  = split(relu6(dw3x3(6 Hs) + bdw))          (CUDA)
//   GEMM2  D2[t] (128 x COUT_P, TMEM) += A2[t] * W3c^T                                      (tensor)
// and finally EPI2: out = s3 * D2 + b3 (+ x).  GEMM1 of chunk c+1 and GEMM2 of chunk c run on the
// tensor pipe while the worker warps do EPI1/DW; the depthwise phase (shared-memory loads) is the
// critical path.  CTAs are persistent (grid = #SMs).  Weights (fp16 hi/lo, split-16x3 scheme of
// kernels_tc.cuh) are either resident in shared memory for the whole kernel (early blocks, <= 83 KB) or
// streamed chunk by chunk through a 2-3-slot bulk-copy (TMA) ring (late blocks).
//
// Roles: warps 0..NWW-1 = workers (thread = GEMM row / pixel / TMEM lane; channel groups own the column
// octets of a chunk), warp NWW = MMA issuer + weight loader (converged warp, issue under elect.sync).
// smem operand tiles use the canonical K-major no-swizzle layout of tc_common.cuh (SBO 128 B,
// LBO = rows/8 * 128 B).  DESIGN.md section 5 lists the measurements behind each of these choices.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace syn {

constexpr int ceil_div_c(int a, int b) { return (a + b - 1) / b; }
// tile shapes / chunk widths / batching depths worth re-measuring when the kernel changes: build variants
// with -D... and compare them on one box with scripts/ab_variants.sh
#ifndef SYN_RO_STEM
#define SYN_RO_STEM 6
#endif
#ifndef SYN_RO_B2
#define SYN_RO_B2 6
#endif
#ifndef SYN_PREP_BATCH
#define SYN_PREP_BATCH 4
#endif
#ifndef SYN_NC_B4
#define SYN_NC_B4 16
#endif
#ifndef SYN_NC_B2
#define SYN_NC_B2 32
#endif
// block 3: 48-channel chunks on 10-row strips (3 chunks per tile instead of 9 with 16 channels on 15 rows: fewer
// per-chunk barrier / wait round trips) measured 0.307 vs 0.325 ms; 48 channels on 6-row strips 0.381 ms
#ifndef SYN_NC_B3
#define SYN_NC_B3 48
#endif
#ifndef SYN_RO_B3
#define SYN_RO_B3 10
#endif
#ifndef SYN_RO_B4
#define SYN_RO_B4 15
#endif
// programmatic dependent launch over the fused launches (measured -1.1 % of the step, round 2)
#ifndef SYN_PDL
#define SYN_PDL 1
#endif
#ifndef SYN_DW2_SMALL
#define SYN_DW2_SMALL 1
#endif
// SYN_DW3: depthwise items of one channel quad x 2 output rows x 5 output columns on the stride-1 60^2 / 30^2 / 15^2 maps
// (15.2 shared-memory loads per 16 outputs instead of 26 + conflicted tap loads)
#ifndef SYN_DW3
#define SYN_DW3 0      // measured SLOWER than the 2 x 2 items (2.46 vs 2.32 ms/step): one long item per thread serialises
#endif
// output columns of a DW3 item: 3 (76 live registers; 17 warps leave 96 per thread: 5 warps share one sub-partition's
// 16 K registers) or 5 (fewer loads per output, needs ~105 registers: spills unless the CTA has <= 16 warps)
#ifndef SYN_DW3_S
#define SYN_DW3_S 3
#endif
#ifndef SYN_DW2_MAXW
#define SYN_DW2_MAXW 30
#endif
#ifndef SYN_NC_B56
#define SYN_NC_B56 64
#endif
#ifndef SYN_NC_B7
#define SYN_NC_B7 64
#endif
#ifndef SYN_NC_B12
#define SYN_NC_B12 64
#endif
#ifndef SYN_NC_B14
#define SYN_NC_B14 64
#endif
#ifndef SYN_NC_B17
#define SYN_NC_B17 32
#endif
#ifndef SYN_EB_STEM
#define SYN_EB_STEM 1
#endif
#ifndef SYN_EB_WIDE
#define SYN_EB_WIDE 1
#endif
#ifndef SYN_EB_MID
#define SYN_EB_MID 1
#endif
#ifndef SYN_EB_SMALL
#define SYN_EB_SMALL 1
#endif
constexpr int round_up_c(int a, int b) { return ceil_div_c(a, b) * b; }
constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int CIN_, int CHID_, int NC_, int COUT_, int W_, int STRIDE_, int RO_, int FACES_, bool RES_, bool STEM_,
          int WSTREAM_>
struct FusedCfg {
  static constexpr bool STEM = STEM_;            // GEMM1 = im2col(3x3 s2 stem conv), CIN = 27 taps
  static constexpr bool RES = RES_;
  static constexpr bool WSTREAM = WSTREAM_ > 0;  // weights streamed per chunk (ring of WSTREAM_ slots) instead of resident
  static constexpr int CIN = CIN_;               // channels of the NHWC input (STEM: 27)
  static constexpr int CIN_P = round_up_c(CIN_, 16);
  static constexpr int CHID = CHID_, NC = NC_, NCHUNK = CHID_ / NC_;
  static constexpr int COUT = COUT_, COUT_P = round_up_c(COUT_, 16);
  static constexpr int NSPLIT = ceil_div_c(COUT_P, 256);                 // MMA N <= 256
  static constexpr int N2 = COUT_P / NSPLIT;
  static constexpr int W = W_, STRIDE = STRIDE_, WO = (W_ - 1) / STRIDE_ + 1;
  static constexpr int RO = RO_, STRIPS = WO / RO_, FACES = FACES_;
  static constexpr int RWIN = (RO_ - 1) * STRIDE_ + 3;                   // window rows incl. halo
  static constexpr int ROWS_MAX = (RWIN < W_ ? RWIN : W_);               // valid input rows per face
  static constexpr int M1_MAX = FACES_ * ROWS_MAX * W_;
  static constexpr int MT1 = ceil_div_c(M1_MAX, 128);
  // DW3 (register-blocked 2 x 5 depthwise items, lanes = NS column segments x NR row pairs): the lane mapping is
  // bank-conflict free only for certain row pitches of the hidden window (HS_COLS) and of the GEMM2 operand (WOP)
  static constexpr int DW3_S = SYN_DW3_S;                                // output columns per item (odd)
  static constexpr bool DW3 = SYN_DW3 && STRIDE_ == 1 && (WO % 15 == 0);                  // 60, 30, 15
  static constexpr int DW3_NS = (WO >= 60) ? 4 : 2, DW3_NR = 8 / DW3_NS;
  // pitch of an output row in the GEMM2 M index (pad columns are computed by the MMA and never read)
  static constexpr int WOP = !DW3 ? WO : (DW3_NS == 4 ? WO + ((6 - WO % 4) % 4) : WO | 1);
  static constexpr int M2F = RO_ * WOP;                                  // GEMM2 rows per face
  static constexpr int M2_MAX = FACES_ * M2F;
  static constexpr int MT2 = ceil_div_c(M2_MAX, 128);
  static constexpr int HS_COLS = (DW3 && DW3_NS == 2) ? ((W_ + 2) | 1) : W_ + 2;      // DW3 with 4 row-pair lanes: odd pitch
  static constexpr int HS_FACE = RWIN * HS_COLS, HS_PIX = FACES_ * HS_FACE, HS_STRIDE = NC_ + 4;
  static constexpr int DWS = NC_ + 4;     // floats between the tap rows of a chunk: mirrored lanes of the 2x2 depthwise read
                                          // taps kx and 2-kx, which must not lie a multiple of 128 bytes apart
  static constexpr int D2_COL = round_up_c(MT1 * NC_, 32);
  // EPI1 TMEM loads kept in flight per wait (measured per map size, scripts/ab_variants.sh)
  static constexpr int EPI1_BATCH = STEM_ ? SYN_EB_STEM : W_ >= 30 ? SYN_EB_WIDE : W_ >= 15 ? SYN_EB_MID : SYN_EB_SMALL;
  // GEMM1's A operand (the block input as fp16 hi/lo) lives in TMEM, not in shared memory: per M tile
  // CIN_P/2 columns of hi K-pairs, then CIN_P/2 columns of lo K-pairs
  static constexpr int XA_COL = D2_COL + MT2 * COUT_P;
  static constexpr int TM_COLS = pow2_cols(XA_COL + MT1 * CIN_P);
  // ---- weight image: [b3 | s3] then NCHUNK x { W1c hi, W1c lo, W3c hi, W3c lo, DW rows } -----------
  static constexpr int B3_BYTES = round_up_c(2 * COUT_P * 4, 128);       // [2][COUT_P] fp32: b3, s3
  static constexpr int W1_PLANE = NC_ * CIN_P * 2;                       // bytes, one plane of one chunk
  static constexpr int W3_PLANE = COUT_P * NC_ * 2;
  static constexpr int DW_ROWS = 12;   // 9 taps, depthwise bias, expand bias b1, expand output scale s1
  static constexpr int CH_W1 = 0, CH_W3 = 2 * W1_PLANE, CH_DW = CH_W3 + 2 * W3_PLANE;
  static constexpr int CHUNK_BYTES = round_up_c(CH_DW + DW_ROWS * DWS * 4, 128);
  static constexpr int W_BYTES = B3_BYTES + NCHUNK * CHUNK_BYTES;
  static constexpr int WSTAGES = WSTREAM_ > 0 ? WSTREAM_ : NCHUNK;       // chunk slots held in smem
  // ---- shared memory carve-up --------------------------------------------------------------------
  static constexpr int A2_PLANE = MT2 * 128 * NC_ * 2;
  static constexpr int S_B3 = 0;
  static constexpr int S_WCH = S_B3 + B3_BYTES;
  static constexpr int S_X = S_WCH + WSTAGES * CHUNK_BYTES;
  static constexpr int S_A2 = S_X;
  static constexpr int S_H = S_A2 + 2 * A2_PLANE;
  // stem only: staged input rows [3][IN_ROWS][120] fp32, exactly as they lie in the NCHW crop (one bulk
  // copy per channel); the left zero-pad column is a predicate in the im2col gather
  static constexpr int IN_ROWS = 2 * ROWS_MAX + 1, IN_STRIDE = 120;
  static constexpr int S_IN = S_H + HS_PIX * HS_STRIDE * 4;
  static constexpr int S_TOTAL = S_IN + (STEM_ ? 3 * IN_ROWS * IN_STRIDE * 4 : 0);
  static constexpr int SMEM_BYTES = S_TOTAL + 1024;                      // + alignment slack
  static_assert(CHID_ % NC_ == 0 && NC_ % 16 == 0, "hidden chunking");
  static_assert(WO % RO_ == 0, "strips must tile the output");
  static_assert(FACES_ == 1 || RO_ == WO, "multi-face tiles hold whole faces");
  static_assert(XA_COL + MT1 * CIN_P <= 512, "TMEM columns");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
  static_assert(N2 % 16 == 0 && N2 <= 256, "MMA N");
  static_assert(WSTREAM_ == 0 || (WSTREAM_ >= 2 && WSTREAM_ <= 4 && NCHUNK >= WSTREAM_), "weight ring");
  // DW3 bank-conflict conditions.  Window loads: lane (s, r) of a quarter-warp reads 16 bytes at group offset
  // S*G*s + 2*HS_COLS*G*r (S = DW3_S, G = HS_STRIDE/4, both odd) and the eight offsets must differ mod 8; operand stores: 8-byte halves
  // at 16-byte slot (S*s + 2*WOP*r) mod 8.  NS=4,NR=2 needs the row-pair term = 4 (mod 8), NS=2,NR=4 needs it = 2 or 6.
  static_assert(!DW3 || ((HS_STRIDE / 4) % 2 == 1), "pixel pitch must be an odd number of 16-byte groups");
  static_assert(!DW3 || (DW3_NS == 4 ? (2 * HS_COLS * (HS_STRIDE / 4)) % 8 == 4 : (2 * HS_COLS * (HS_STRIDE / 4)) % 4 == 2), "window row pitch");
  static_assert(!DW3 || (DW3_NS == 4 ? (2 * WOP) % 8 == 4 : (2 * WOP) % 4 == 2), "operand row pitch");
  static_assert(!DW3 || FACES_ == 1 || WOP == WO, "padded output rows only with one face per tile");
};

struct FusedArgs {
  const float* x;        // NHWC (B,W,W,CIN) -- or NCHW (B,3,120,120) for the stem variant
  const uint8_t* x_u8;   // stem variant only: raw uint8 crop, normalised (v-127.5)/128 while staging; else null
  const uint8_t* wimg;   // packed weight image (FusedCfg::W_BYTES)
  float* y;              // NHWC (B,WO,WO,COUT)
  int batch;
  int split;             // two-face configs: face groups >= split hold ONE face (tail wave, see fused_tile_plan)
  int face_groups;       // number of face groups (tiles = face_groups * STRIPS)
  int* err;              // sticky time-out flag of the bounded mbarrier waits (mapped pinned host memory)
  int* sat;              // sticky "a block input was clamped to the fp16 range" flag (device memory)
  int border;            // uint8 stem only: CenterCrop margin, pixels of the frame read as 0 (utils/ddfa.py:162-243); 0 = off
  int npass;             // 3 = split-fp16 x3 (hi*hi + hi*lo + lo*hi); 1 = single fp16 pass (SYN_ENGINE_TC_FUSED_1PASS)
#ifdef SYN_FUSED_TRACE
  int trace_id;          // backbone block of this launch (1..17)
#endif
};

// Phase trace (debug builds only, -DSYN_FUSED_TRACE): clock64 stamps of CTA 0's second tile, one row of 8
// events per (block, role, chunk); read back with syn_debug_read_trace.  role 0 = worker thread 0, 1 = issuer.
#ifdef SYN_FUSED_TRACE
__device__ long long g_fused_trace[18 * 2 * 64 * 8];
#define SYN_TRACE(role, chunk, ev)                                                                        \
  do {                                                                                                    \
    if (trace_on) g_fused_trace[((p.trace_id * 2 + (role)) * 64 + (chunk)) * 8 + (ev)] = clock64();      \
  } while (0)
#else
#define SYN_TRACE(role, chunk, ev) do { } while (0)
#endif

// Named barrier of a worker group with an IMMEDIATE id (a register id makes ptxas reserve all 16 hardware barriers).
template <int THREADS>
__device__ __forceinline__ void group_bar_sync(int grp) {
  switch (grp) {
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(THREADS) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"n"(THREADS) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(THREADS) : "memory"); break;
  }
}

// NWW = worker warps (multiple of 4: TMEM lane quarter = warp % 4); the issuer is warp NWW.
template <class C, int NWW>
__global__ void __launch_bounds__((NWW + 1) * 32, 1) fused_mbconv_kernel(const FusedArgs p) {
  constexpr int NWT = NWW * 32;          // worker threads
  constexpr int NWG = NWW / 4;           // worker groups: group g owns every NWG-th (tile, column-chunk) pair
  static_assert(NWW % 4 == 0 && NWW >= 4 && NWW <= 24, "worker warps");
  // Channel groups: the hidden channels of a chunk are split between NG groups of worker warps.  A group
  // drains ITS channels from TMEM (EPI1) and runs the depthwise conv on ITS channels, so the only
  // synchronisation between EPI1 and DW is a named barrier among the group's warps, and the groups drift
  // freely against each other (one can be in EPI1 while another is in DW).
  constexpr int NKG_ = C::NC / 8;
  constexpr int NG = (NWG % 4 == 0 && NKG_ % 4 == 0) ? 4 : (NWG % 2 == 0 && NKG_ % 2 == 0) ? 2 : 1;
  constexpr int WPG = NWW / NG, TPG = WPG * 32;      // warps / threads per group (WPG is a multiple of 4)
  constexpr int KPG = NKG_ / NG;                       // 8-channel groups owned by a worker group
  constexpr int SUBS = WPG / 4;                        // sub-groups of 128 threads (one TMEM lane each)
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_wfull[4], bar_x, bar_d1, bar_epi1, bar_a2, bar_g2, bar_d2free, bar_in;
  __shared__ uint32_t tmem_base_s;

  // keep the pointer in the shared address space (no integer round trip): a generic pointer here
  // turns every tile access into LD.E/ST.E instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & 127, wg = tid >> 7;   // GEMM row / TMEM lane of this worker, and its 128-thread slice
  const int grp = warp / WPG;                  // channel group (workers only)
  const int gtid = tid - grp * TPG, gsub = gtid >> 7;
  const int ntiles = p.face_groups * C::STRIPS;
  // faces of a face group.  Two-face tiles (8x8 maps): 512 groups over 148 SMs would leave the last wave
  // 46 % full, so the host turns the groups of that wave into single-face groups (twice as many CTAs busy,
  // each done sooner); small batches become single-face groups altogether.
  auto group_faces = [&](int fg, int& f0, int& nfaces) {
    if constexpr (C::FACES == 2) {
      if (fg < p.split) { f0 = 2 * fg; nfaces = 2; }
      else { f0 = 2 * p.split + (fg - p.split); nfaces = 1; }
    } else {
      f0 = fg * C::FACES;
      nfaces = min(C::FACES, p.batch - f0);
    }
  };

#if SYN_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel's prologue may overlap this one's tail
#endif
  if (tid == 0) {
    mbar_init(smem_u32(&bar_w), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar_wfull[i]), 1);
    mbar_init(smem_u32(&bar_x), NWT);
    mbar_init(smem_u32(&bar_d1), 1);
    mbar_init(smem_u32(&bar_epi1), NWT);
    mbar_init(smem_u32(&bar_a2), NWT);
    mbar_init(smem_u32(&bar_g2), 1);
    mbar_init(smem_u32(&bar_d2free), NWT);
    mbar_init(smem_u32(&bar_in), 1);
    fence_mbar_init();
  }
  if (warp == NWW) tmem_alloc<C::TM_COLS>(smem_u32(&tmem_base_s));
  if constexpr (C::STEM) {     // staged-row buffer: the left pad column (and everything else) starts as zero;
    float* z = reinterpret_cast<float*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u) + C::S_IN);
    for (int i = threadIdx.x; i < 3 * C::IN_ROWS * C::IN_STRIDE / 4; i += blockDim.x)   // before any bulk copy
      reinterpret_cast<float4*>(z)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  uint8_t* sWch = smem + C::S_WCH;
  uint8_t* sA2 = smem + C::S_A2;
  float* sH = reinterpret_cast<float*>(smem + C::S_H);
  const float* sB3 = reinterpret_cast<const float*>(smem + C::S_B3);
  float* sIn = reinterpret_cast<float*>(smem + C::S_IN);   // stem variant only

  if (warp < NWW) {
    // =============================== workers ====================================================
    // zero the whole hidden window once: halo columns are never written afterwards
    for (int i = tid; i < C::HS_PIX * C::HS_STRIDE / 4; i += NWT)
      reinterpret_cast<float4*>(sH)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_wait(smem_u32(&bar_w), 0, p.err);                // b3/s3 (and, if resident, all chunks) landed
    uint32_t n_d1 = 0, n_g2 = 0, g = 0, n_in = 0;                   // completed-phase counters; g = chunk counter
    asm volatile("bar.sync 5, %0;" ::"n"(NWT) : "memory");

    // Geometry of a tile + "prep": stage / convert its input into the GEMM1 A operand and publish it.
    // prep(next tile) is issued BEFORE the current tile's EPI2, so the issuer can run GEMM1(next, 0) --
    // and the global-load latency of the conversion is hidden -- while the workers drain D2.
    auto prep = [&](int tile) {
      const int fg = tile / C::STRIPS, sp = tile - fg * C::STRIPS;
      int f0, nfaces;
      group_faces(fg, f0, nfaces);
      const int iy0 = sp * C::RO * C::STRIDE - 1;
      const int rf = max(iy0, 0), rl = min(iy0 + C::RWIN - 1, C::W - 1);
      const int ppf = (rl - rf + 1) * C::W;
      const int M1 = nfaces * ppf;
      const int mt1 = (M1 + 127) >> 7;
#ifdef SYN_FUSED_TRACE
      const bool trace_on = blockIdx.x == 0 && tile == 2 * (int)gridDim.x && tid == 0;   // prep of the tile after the traced one
#endif
      SYN_TRACE(0, 62, 0);
      // ---- stem: the crop rows this strip needs, zero outside the image ------------------------------
      if constexpr (C::STEM) {
        const int iy_first = 2 * rf - 1, nin = 2 * (rl - rf + 1) + 1;
        if (p.x_u8 != nullptr) {              // uint8 crops: threads load, normalise and stage
          for (int i = tid; i < 3 * nin * 30; i += NWT) {
            const int c4 = i % 30, r = (i / 30) % nin, ci = i / (30 * nin);
            const int iy = iy_first + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < kImg) {
              uchar4 u = *reinterpret_cast<const uchar4*>(p.x_u8 + ((size_t)(f0 * 3 + ci) * kImg + iy) * kImg + c4 * 4);
              if (p.border > 0) {                                   // zero frame of the reference loader, before normalisation
                const int col = c4 * 4;
                const bool row_out = iy < p.border || iy >= kImg - p.border;
                if (row_out || col < p.border || col >= kImg - p.border) u.x = 0;
                if (row_out || col + 1 < p.border || col + 1 >= kImg - p.border) u.y = 0;
                if (row_out || col + 2 < p.border || col + 2 >= kImg - p.border) u.z = 0;
                if (row_out || col + 3 < p.border || col + 3 >= kImg - p.border) u.w = 0;
              }
              v = make_float4(((float)u.x - 127.5f) / 128.0f, ((float)u.y - 127.5f) / 128.0f,
                              ((float)u.z - 127.5f) / 128.0f, ((float)u.w - 127.5f) / 128.0f);
            }
            *reinterpret_cast<float4*>(sIn + (ci * C::IN_ROWS + r) * C::IN_STRIDE + c4 * 4) = v;
          }
        } else {                              // fp32 crops: rows were bulk-copied by the issuer one tile ahead
          mbar_wait(smem_u32(&bar_in), n_in & 1, p.err);
          ++n_in;
          for (int r = 0; r < nin; ++r) {     // rows outside the image are not copied: zero them (edge strips)
            const int iy = iy_first + r;
            if (iy < 0 || iy >= kImg)
              for (int i = tid; i < 3 * 30; i += NWT)
                *reinterpret_cast<float4*>(sIn + ((i / 30) * C::IN_ROWS + r) * C::IN_STRIDE + (i % 30) * 4) =
                    make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        SYN_TRACE(0, 62, 1);
        asm volatile("bar.sync 5, %0;" ::"n"(NWT) : "memory");
      }
      SYN_TRACE(0, 62, 2);
      // ---- X tile -> fp16 hi/lo K pairs in TMEM ---------------------------------------------------
      // The conversion sits between the last depthwise and EPI2 of the current tile, so its global-load
      // latency is exposed once per batch of loads: all loads of a batch are issued unconditionally (from a
      // clamped, always valid address) before the first use, and masked afterwards -- a predicated load per
      // item would put a branch between the loads and serialise one DRAM latency per item.
      {
        constexpr int KG = C::CIN_P / 8;
        constexpr int ITERS = (C::MT1 * KG + NWG - 1) / NWG;             // items per thread
        constexpr int PB = ITERS < SYN_PREP_BATCH ? ITERS : SYN_PREP_BATCH;
        const int n_items = mt1 * KG;
        for (int e0 = wg; e0 < n_items; e0 += PB * NWG) {
          float v[PB][8];
          if constexpr (C::STEM) {
            // im2col of the 3x3 stride-2 pad-1 stem conv from the staged rows: k = (ci*3+ky)*3+kx.  All PB items of a
            // thread are gathered before the first conversion (independent shared loads in flight instead of one
            // load -> convert -> TMEM-store chain per item); the kg switch makes every tap offset a compile-time constant.
#pragma unroll
            for (int u = 0; u < PB; ++u) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
              const int e = e0 + u * NWG;
              if (e >= n_items) continue;
              const int t = e / KG, kg = e - t * KG;
              const int mr = t * 128 + row;                     // FACES == 1
              if (mr < M1) {
                const int yl = mr / C::W, xx = mr - yl * C::W;
                const float* base = sIn + (2 * yl) * C::IN_STRIDE + 2 * xx - 1;   // column 2xx-1+kx; -1 is the zero pad
#pragma unroll
                for (int kgc = 0; kgc < KG; ++kgc)
                  if (kg == kgc) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const int k = kgc * 8 + j;
                      if (k < 27) {
                        const int ci = k / 9, ky = (k % 9) / 3, kx = k % 3;
                        if (kx > 0 || xx > 0) v[u][j] = base[(ci * C::IN_ROWS + ky) * C::IN_STRIDE + kx];
                      }
                    }
                  }
              }
            }
          } else {
            float4 qa[PB], qb[PB];
            bool ok[PB];
#pragma unroll
            for (int u = 0; u < PB; ++u) {
              const int e = min(e0 + u * NWG, n_items - 1);
              const int t = e / KG, kg = e - t * KG;
              const int m = t * 128 + row;
              ok[u] = (e0 + u * NWG < n_items) && (m < M1) && (kg * 8 < C::CIN);
              const int mc = min(m, M1 - 1), kgc = min(kg, (C::CIN - 1) / 8);
              const int f = (C::FACES > 1) ? mc / ppf : 0;
              const int mr = mc - f * ppf;                    // pixel inside the face's valid rows
              const float* src = p.x + ((size_t)((f0 + f) * C::W + rf) * C::W + mr) * C::CIN + kgc * 8;
              qa[u] = __ldg(reinterpret_cast<const float4*>(src));
              qb[u] = __ldg(reinterpret_cast<const float4*>(src + 4));
            }
            if (e0 == wg) SYN_TRACE(0, 62, 5);
#pragma unroll
            for (int u = 0; u < PB; ++u) {
              v[u][0] = ok[u] ? qa[u].x : 0.f; v[u][1] = ok[u] ? qa[u].y : 0.f; v[u][2] = ok[u] ? qa[u].z : 0.f; v[u][3] = ok[u] ? qa[u].w : 0.f;
              v[u][4] = ok[u] ? qb[u].x : 0.f; v[u][5] = ok[u] ? qb[u].y : 0.f; v[u][6] = ok[u] ? qb[u].z : 0.f; v[u][7] = ok[u] ? qb[u].w : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < PB; ++u) {
            const int e = e0 + u * NWG;
            if (e < n_items) {                                // warp-uniform: tcgen05.st is .sync.aligned
              const int t = e / KG, kg = e - t * KG;
              uint32_t h[4], l[4];
              float vmax = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) vmax = fmaxf(vmax, fabsf(v[u][j]));
              if (vmax * kActScale > 60000.f) *p.sat = 1;          // the clamp below changes a value: tell the host (sticky)
#pragma unroll
              for (int j = 0; j < 4; ++j) split2_f16(v[u][2 * j] * kActScale, v[u][2 * j + 1] * kActScale, h[j], l[j]);
              // TMEM lane = GEMM row of this thread; 8 K values = 4 columns of fp16 pairs
              const uint32_t xa = tmem + ((uint32_t)((warp & 3) * 32) << 16) + C::XA_COL + t * C::CIN_P + kg * 4;
              tmem_st4(xa, h[0], h[1], h[2], h[3]);
              tmem_st4(xa + C::CIN_P / 2, l[0], l[1], l[2], l[3]);
              if (e == wg) SYN_TRACE(0, 62, 6);
            }
          }
          if (e0 == wg) SYN_TRACE(0, 62, 7);
        }
      }
      SYN_TRACE(0, 62, 3);
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&bar_x));
      SYN_TRACE(0, 62, 4);

    };
    // L2 prefetch of a tile's input (one contiguous NHWC range) a whole tile ahead of its conversion: the
    // loads in prep() then hit L2 with warm TLB entries instead of paying ~2000 cycles per batch
    auto prefetch_x = [&](int tile) {
      if constexpr (!C::STEM) {
        if (tile >= ntiles) return;
        const int fg = tile / C::STRIPS, sp = tile - fg * C::STRIPS;
        int f0, nfaces;
        group_faces(fg, f0, nfaces);
        const int iy0 = sp * C::RO * C::STRIDE - 1;
        const int rf = max(iy0, 0), rl = min(iy0 + C::RWIN - 1, C::W - 1);
        const char* base = reinterpret_cast<const char*>(p.x + ((size_t)(f0 * C::W + rf) * C::W) * C::CIN);
        const int bytes = (C::FACES > 1 ? nfaces * C::W * C::W : (rl - rf + 1) * C::W) * C::CIN * 4;
        for (int o = tid * 128; o < bytes; o += NWT * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    };
#if SYN_PDL
    // Programmatic dependent launch: everything above (barriers, TMEM, the zeroed window, the weight image) does not
    // depend on the previous kernel; its output -- this kernel's input -- is first touched below, and this kernel's
    // first global store comes later still.
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
    if ((int)blockIdx.x < ntiles) prep(blockIdx.x);
    prefetch_x(blockIdx.x + gridDim.x);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int fg = tile / C::STRIPS, sp = tile - fg * C::STRIPS;
      int f0, nfaces;
      group_faces(fg, f0, nfaces);
      const int oy0 = sp * C::RO;
      const int iy0 = oy0 * C::STRIDE - 1;
      const int rf = max(iy0, 0), rl = min(iy0 + C::RWIN - 1, C::W - 1);
      const int ppf = (rl - rf + 1) * C::W;               // valid input pixels per face
#ifdef SYN_FUSED_TRACE
      const bool trace_on = blockIdx.x == 0 && tile == (int)gridDim.x && tid == 0;
#endif
      SYN_TRACE(0, 63, 0);
      const int M1 = nfaces * ppf;
      const int mt1 = (M1 + 127) >> 7;
      const int M2 = nfaces * C::M2F;
      const int mt2 = (M2 + 127) >> 7;

      for (int c = 0; c < C::NCHUNK; ++c, ++g) {
        const int slot = C::WSTREAM ? (int)(g % C::WSTAGES) : c;
        if constexpr (C::WSTREAM) mbar_wait(smem_u32(&bar_wfull[slot]), (g / C::WSTAGES) & 1, p.err);
        const float* dwc = reinterpret_cast<const float*>(sWch + slot * C::CHUNK_BYTES + C::CH_DW);
        SYN_TRACE(0, c, 0);
        // ---- EPI1: D1 -> relu6(s1*D1 + b1) -> hidden window --------------------------------------
        mbar_wait(smem_u32(&bar_d1), n_d1 & 1, p.err);
        ++n_d1;
        tc_fence_after_sync();
        SYN_TRACE(0, c, 1);
        group_bar_sync<TPG>(grp);   // the group is done reading ITS Hs columns (DW c-1)
        SYN_TRACE(0, c, 2);
        if (c == 0) {
          // ---- strip mode: window rows outside the image must read as zero (may hold a previous tile);
          //      every group clears its own channel columns
          if constexpr (C::STRIPS > 1) {
            constexpr int CQ = KPG * 2;                               // float4 per pixel owned by the group
            if (iy0 < 0)
              for (int i = gtid; i < C::HS_COLS * CQ; i += TPG)
                *reinterpret_cast<float4*>(sH + (size_t)(i / CQ) * C::HS_STRIDE + grp * KPG * 8 + (i % CQ) * 4) =
                    make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy0 + C::RWIN - 1 > C::W - 1)
              for (int i = gtid; i < C::HS_COLS * CQ; i += TPG)
                *reinterpret_cast<float4*>(sH + (size_t)((C::RWIN - 1) * C::HS_COLS + i / CQ) * C::HS_STRIDE +
                                           grp * KPG * 8 + (i % CQ) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        {
          // EPI1: 8 columns (one channel octet) per TMEM load; the expand scale is one power of two per
          // layer (row 11 is constant) and the octet's biases stay in registers, so an element costs one
          // FFMA.SAT plus a quarter of a 16-byte shared store
          const float sc1 = dwc[11 * C::DWS];
          int cur_k = -1;
          float bq[8];
          const int n_e = mt1 * KPG;
          constexpr int EB = C::EPI1_BATCH;
          for (int e0 = gsub; e0 < n_e; e0 += EB * SUBS) {
            uint32_t vr[EB][8];
#pragma unroll
            for (int u = 0; u < EB; ++u) {                  // up to EB TMEM loads in flight, one wait
              const int e = e0 + u * SUBS;
              if (e < n_e) {                                // warp-uniform
                const int t = e / KPG, kq = grp * KPG + (e - t * KPG);
                tmem_ld8_async(tmem + ((uint32_t)((warp & 3) * 32) << 16) + t * C::NC + kq * 8, vr[u]);
              }
            }
            tmem_wait_ld();
#pragma unroll
            for (int u = 0; u < EB; ++u) {
              const int e = e0 + u * SUBS;
              if (e < n_e) {
                const int t = e / KPG, kq = grp * KPG + (e - t * KPG), j0 = kq * 8;
                if (kq != cur_k) {
                  cur_k = kq;
                  const float4 b0 = *reinterpret_cast<const float4*>(dwc + 10 * C::DWS + j0);
                  const float4 b1 = *reinterpret_cast<const float4*>(dwc + 10 * C::DWS + j0 + 4);
                  bq[0] = b0.x; bq[1] = b0.y; bq[2] = b0.z; bq[3] = b0.w; bq[4] = b1.x; bq[5] = b1.y; bq[6] = b1.z; bq[7] = b1.w;
                }
                const int m = t * 128 + row;
                if (m < M1) {
                  const int f = (C::FACES > 1) ? m / ppf : 0;
                  const int mr = m - f * ppf;
                  const int yl = mr / C::W, xx = mr - yl * C::W;
                  float* hrow = sH + (size_t)(f * C::HS_FACE + (rf - iy0 + yl) * C::HS_COLS + xx + 1) * C::HS_STRIDE + j0;
                  *reinterpret_cast<float4*>(hrow) =
                      make_float4(__saturatef(fmaf(__uint_as_float(vr[u][0]), sc1, bq[0])), __saturatef(fmaf(__uint_as_float(vr[u][1]), sc1, bq[1])),
                                  __saturatef(fmaf(__uint_as_float(vr[u][2]), sc1, bq[2])), __saturatef(fmaf(__uint_as_float(vr[u][3]), sc1, bq[3])));
                  *reinterpret_cast<float4*>(hrow + 4) =
                      make_float4(__saturatef(fmaf(__uint_as_float(vr[u][4]), sc1, bq[4])), __saturatef(fmaf(__uint_as_float(vr[u][5]), sc1, bq[5])),
                                  __saturatef(fmaf(__uint_as_float(vr[u][6]), sc1, bq[6])), __saturatef(fmaf(__uint_as_float(vr[u][7]), sc1, bq[7])));
                }
              }
            }
          }
        }
        SYN_TRACE(0, c, 3);
        tc_fence_before_sync();
        mbar_arrive(smem_u32(&bar_epi1));
        group_bar_sync<TPG>(grp);   // the group's channel columns of the window are complete
        // ---- DW: 3x3 depthwise on the window -> A2 operand ----------------------------------------
        if (c > 0) {                                        // A2 is free once GEMM2(c-1) has completed
          mbar_wait(smem_u32(&bar_g2), n_g2 & 1, p.err);
          ++n_g2;
        }
        SYN_TRACE(0, c, 4);
        if constexpr (C::DW3) {
          // Stride-1 60^2 / 30^2 / 15^2 maps.  The depthwise phase is bound by shared-memory wavefronts (every LDS.128 of
          // a warp costs four), so an item is register-blocked as far as the register file allows: ONE channel quad x
          // TWO output rows x S = 3 (5) output columns = 4 x (S + 2) window loads + 10 tap loads per 8 S outputs: 20 (15.2)
          // per 16 outputs; the 2 x 2 items below need 26 and their mirrored tap loads used to conflict.  A unit of 16
          // threads = 8 lanes (NS column segments x NR row pairs) x the two quads of a channel octet.  Segments
          // are S pixels = an odd number of 16-byte groups apart, row pairs 2 * HS_COLS pixels: with the row pitches
          // FusedCfg asserts, the eight window addresses of a quarter-warp and the sixteen 8-byte operand stores of a
          // half-warp fall on different banks.
          constexpr int NS = C::DW3_NS, NR = C::DW3_NR, S = C::DW3_S, NSEG = C::WO / S, RP2 = (C::RO + 1) / 2;
          constexpr int XGU = (NSEG + NS - 1) / NS, RPGU = (RP2 + NR - 1) / NR, UPK = XGU * RPGU;   // units per (face, octet)
          const int l8 = tid & 7, qh = (tid >> 3) & 1;
          const int ls = l8 % NS, lr = l8 / NS;
          const int units = KPG * nfaces * UPK;
          for (int u = gtid >> 4; u < units; u += TPG / 16) {
            const int kgl = u / (nfaces * UPK), r1 = u - kgl * (nfaces * UPK);
            const int f = r1 / UPK, r2 = r1 - f * UPK;
            const int rpg = r2 / XGU, xg = r2 - rpg * XGU;
            const int kg = grp * KPG + kgl, j0 = kg * 8 + qh * 4;
            const int seg = xg * NS + ls, oy = 2 * (rpg * NR + lr);
            if (seg >= NSEG || oy >= C::RO) continue;
            const int ox0 = S * seg;
            const bool two = (oy + 1 < C::RO);
            const float* wq = dwc + j0;
            float2 w[3][3][2];                                             // [ky][kx][channel pair]
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const float4 t4 = *reinterpret_cast<const float4*>(wq + (ky * 3 + kx) * C::DWS);
                w[ky][kx][0] = make_float2(t4.x, t4.y); w[ky][kx][1] = make_float2(t4.z, t4.w);
              }
            float2 acc[2][S][2];                                           // [output row][output column][channel pair]
            {
              const float4 b4 = *reinterpret_cast<const float4*>(wq + 9 * C::DWS);
#pragma unroll
              for (int ro = 0; ro < 2; ++ro)
#pragma unroll
                for (int a = 0; a < S; ++a) { acc[ro][a][0] = make_float2(b4.x, b4.y); acc[ro][a][1] = make_float2(b4.z, b4.w); }
            }
            // window columns ox0-1 .. ox0+S are Hs columns ox0 .. ox0+S+1; window rows oy .. oy+3
            const float* hb = sH + (size_t)(f * C::HS_FACE + oy * C::HS_COLS + ox0) * C::HS_STRIDE + j0;
#pragma unroll
            for (int ic = 0; ic < S + 2; ++ic) {
              float2 d[4][2];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                if (r == 3 && !two) continue;                              // row only the absent second output row needs
                const float4 t4 = *reinterpret_cast<const float4*>(hb + (r * C::HS_COLS + ic) * C::HS_STRIDE);
                d[r][0] = make_float2(t4.x, t4.y); d[r][1] = make_float2(t4.z, t4.w);
              }
#pragma unroll
              for (int a = 0; a < S; ++a) {
                const int kx = ic - a;
                if (kx < 0 || kx > 2) continue;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) acc[0][a][j] = ffma2(d[ky][j], w[ky][kx][j], acc[0][a][j]);
                  if (two) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) acc[1][a][j] = ffma2(d[ky + 1][j], w[ky][kx][j], acc[1][a][j]);
                  }
                }
              }
            }
            constexpr float kOut = 6.0f * kActScale;
#pragma unroll
            for (int ro = 0; ro < 2; ++ro) {
              if (ro == 1 && !two) continue;
#pragma unroll
              for (int a = 0; a < S; ++a) {
                const int m2 = f * C::M2F + (oy + ro) * C::WOP + ox0 + a;
                uint32_t h0, l0, h1, l1;
                split2_f16<false>(__saturatef(acc[ro][a][0].x) * kOut, __saturatef(acc[ro][a][0].y) * kOut, h0, l0);
                split2_f16<false>(__saturatef(acc[ro][a][1].x) * kOut, __saturatef(acc[ro][a][1].y) * kOut, h1, l1);
                uint8_t* dst = sA2 + (m2 >> 7) * (128 * C::NC * 2) + ((m2 & 127) >> 3) * 128 + kg * 2048 + (m2 & 7) * 16 + qh * 8;
                *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(dst + C::A2_PLANE) = make_uint2(l0, l1);
              }
            }
          }
        } else if constexpr (C::STRIDE == 1 && ((C::WO >= 15 && C::WO <= SYN_DW2_MAXW) || (SYN_DW2_SMALL && C::WO == 8))) {
          // Stride-1 30^2, 15^2 and 8^2 maps (on the 60^2 map of block 1 the units do not divide evenly between
          // the channel groups and the row-pair items below are faster): the window loads of the depthwise
          // conv are what the shared-memory pipe spends its time on, so an item is register-blocked over a 2 x 2 output patch
          // of ONE channel quad -- 16 window + 10 tap LDS.128 per 16 outputs instead of 24 + 20.
          // A unit of 16 threads = 8 lanes along x (column pairs) x the two quads of a channel octet.
          // Lanes 4-7 are MIRRORED: they walk the four window columns right to left, use the taps with kx
          // reversed and own the patch columns in the opposite order.  Neighbouring lanes are 2 pixels =
          // an even number of 16-byte bank groups apart, and the mirror image shifts lanes 4-7 onto the odd
          // groups: every window load (quarter-warp) and every 8-byte operand store (half-warp) is
          // bank-conflict free.
          // 8x8 maps (SYN_DW2_SMALL; measured -2.2 % of the step with scripts/quick_variant_check.py): the 8 lanes
          // are 4 column pairs x 2 row pairs; the second row pair lies 2 window rows = 4 bank groups further
          // and is the mirrored half.
          constexpr int XL = (C::WO >= 15) ? 8 : 4, YL = 8 / XL;           // lanes of a unit along x / along row pairs
          constexpr int CPR = (C::WO + 1) / 2, XG2 = (CPR + XL - 1) / XL, RP2 = (C::RO + 1) / 2;
          constexpr int RPU = (RP2 + YL - 1) / YL, UPK = RPU * XG2;        // units per (face, channel octet)
          const int l8 = tid & 7, qh = (tid >> 3) & 1;
          const int lx = l8 % XL, ly = l8 / XL;
          const bool mir = (l8 & 4) != 0;
          const int units = KPG * nfaces * UPK;
          for (int u = gtid >> 4; u < units; u += TPG / 16) {
            const int kgl = u / (nfaces * UPK), r1 = u - kgl * (nfaces * UPK);
            const int f = r1 / UPK, r2 = r1 - f * UPK;
            const int rpu = r2 / XG2, xg = r2 - rpu * XG2;
            const int kg = grp * KPG + kgl, j0 = kg * 8 + qh * 4;
            const int ox = 2 * (xg * XL + lx), oy = 2 * (rpu * YL + ly);
            if (ox >= C::WO || oy >= C::RO) continue;
            const bool two = (oy + 1 < C::RO);
            const float* wq = dwc + j0;
            float2 w[3][3][2];                                             // [ky][local kx][channel pair] (FFMA2 operands)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kl = 0; kl < 3; ++kl) {
                const float4 t4 = *reinterpret_cast<const float4*>(wq + (ky * 3 + (mir ? 2 - kl : kl)) * C::DWS);
                w[ky][kl][0] = make_float2(t4.x, t4.y); w[ky][kl][1] = make_float2(t4.z, t4.w);
              }
            float2 acc[2][2][2];                                           // [output row][local column][channel pair]
            {
              const float4 b4 = *reinterpret_cast<const float4*>(wq + 9 * C::DWS);
#pragma unroll
              for (int ro = 0; ro < 2; ++ro)
#pragma unroll
                for (int a = 0; a < 2; ++a) { acc[ro][a][0] = make_float2(b4.x, b4.y); acc[ro][a][1] = make_float2(b4.z, b4.w); }
            }
            // window columns ox-1 .. ox+2 are Hs columns ox .. ox+3; local column ic is Hs column ox+ic (ox+3-ic mirrored)
            const float* hb = sH + (size_t)(f * C::HS_FACE + oy * C::HS_COLS + ox + (mir ? 3 : 0)) * C::HS_STRIDE + j0;
            const int cstep = mir ? -C::HS_STRIDE : C::HS_STRIDE;
#pragma unroll
            for (int ic = 0; ic < 4; ++ic) {
              float2 d[4][2];
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                if (r == 3 && !two) continue;                              // row only the absent second output row needs
                const float4 t4 = *reinterpret_cast<const float4*>(hb + r * (C::HS_COLS * C::HS_STRIDE) + ic * cstep);
                d[r][0] = make_float2(t4.x, t4.y); d[r][1] = make_float2(t4.z, t4.w);
              }
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const int kl = ic - a;
                if (kl < 0 || kl > 2) continue;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) acc[0][a][j] = ffma2(d[ky][j], w[ky][kl][j], acc[0][a][j]);
                  if (two) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) acc[1][a][j] = ffma2(d[ky + 1][j], w[ky][kl][j], acc[1][a][j]);
                  }
                }
              }
            }
            constexpr float kOut = 6.0f * kActScale;
#pragma unroll
            for (int ro = 0; ro < 2; ++ro) {
              if (ro == 1 && !two) continue;
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const int col = ox + (mir ? 1 - a : a);
                if (col >= C::WO) continue;                                // odd width: the last pair has one column
                const int m2 = f * C::M2F + (oy + ro) * C::WOP + col;
                uint32_t h0, l0, h1, l1;
                split2_f16<false>(__saturatef(acc[ro][a][0].x) * kOut, __saturatef(acc[ro][a][0].y) * kOut, h0, l0);
                split2_f16<false>(__saturatef(acc[ro][a][1].x) * kOut, __saturatef(acc[ro][a][1].y) * kOut, h1, l1);
                uint8_t* dst = sA2 + (m2 >> 7) * (128 * C::NC * 2) + ((m2 & 127) >> 3) * 128 + kg * 2048 + (m2 & 7) * 16 + qh * 8;
                *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(dst + C::A2_PLANE) = make_uint2(l0, l1);
              }
            }
          }
        } else
        {
          // Item = (8 hidden channels, two vertically adjacent output rows, 8 lanes along x): the 3x3
          // windows of the two rows share (S=1: 2 of 4, S=2: 1 of 5) input rows and all nine tap vectors,
          // which cuts the shared-memory loads per output by ~40 % against one pixel per thread.
          // Hidden values are stored as relu6(h)/6 in [0,1] and the bias row holds bdw/6, so the
          // activation is a single saturate and the fp16 pre-scale becomes 6 * kActScale.
          constexpr int NKG = C::NC / 8;
          constexpr int GX = (C::WO >= 8) ? 8 : 4, GY = 8 / GX;            // quarter-warp footprint
          constexpr int XG = (C::WO + GX - 1) / GX;                        // x groups per output row
          // small maps (8x8, 4x4) have fewer row-pair items than worker threads: one output row per item
          // there, so that every thread has work and the per-chunk dependency chain is half as long
          // (measured: pays when at most a quarter of the threads would have a row-pair item, blocks 7/14/17;
          // with half of them busy the extra window loads of single rows cost more than the idle warps)
          constexpr int RPI = (4 * C::FACES * ((C::RO + 1) / 2) * ((C::WO + GX - 1) / GX) * GX * NKG <= NWT) ? 1 : 2;
          constexpr int RP = (C::RO + RPI - 1) / RPI, RPG = (RP + GY - 1) / GY;   // row pairs (rows), groups of them
          constexpr int PER_FACE = XG * RPG;
          constexpr int NR = (RPI - 1) * C::STRIDE + 3;                    // window rows of an item
          const int per_kg = nfaces * PER_FACE;
          const int l8 = tid & 7, lx = l8 % GX, ly = l8 / GX;
          // Stride 2: the window pixels of neighbouring lanes lie 2 * HS_STRIDE floats apart, an even number
          // of 16-byte bank groups, so lanes l and l+4 of a quarter-warp would collide on every window load.
          // Lanes 4-7 therefore take the two channel quads of their octet in the opposite order (q0, q1 are
          // the float offsets of the first / second quad); only the final operand store swaps them back.
          const bool swz = (C::STRIDE == 2) && (l8 & 4);
          const int q0 = swz ? 4 : 0, q1 = 4 - q0;
          {
            const int kg_end = (grp + 1) * KPG;                              // this group's channel octets
            int kg = grp * KPG, it = gtid >> 3;
            while (it >= per_kg && kg < kg_end) { it -= per_kg; ++kg; }
            while (kg < kg_end) {
              const int f = it / PER_FACE, r2 = it - f * PER_FACE;
              const int rpg = r2 / XG, xg = r2 - rpg * XG;
              // single rows with GY == 2: the two rows of a quarter-warp lie RPG rows apart (an even number),
              // which keeps the two half-rows of lanes on disjoint bank groups
              const int ox = xg * GX + lx, oy = (RPI == 2) ? 2 * (rpg * GY + ly) : rpg + RPG * ly;
              if (ox < C::WO && oy < C::RO) {
                const float* wbase = dwc + kg * 8;
                const float* h0 = sH + (size_t)(f * C::HS_FACE + (oy * C::STRIDE) * C::HS_COLS + ox * C::STRIDE) * C::HS_STRIDE + kg * 8;
                const bool two = (RPI == 2) && (oy + 1 < C::RO);             // second output row exists
                float2 acc0[4], acc1[4];                                     // channel pairs (FFMA2 operands)
                {
                  const float4 a = *reinterpret_cast<const float4*>(wbase + 9 * C::DWS + q0);
                  const float4 e = *reinterpret_cast<const float4*>(wbase + 9 * C::DWS + q1);
                  acc0[0] = make_float2(a.x, a.y); acc0[1] = make_float2(a.z, a.w); acc0[2] = make_float2(e.x, e.y); acc0[3] = make_float2(e.z, e.w);
  #pragma unroll
                  for (int j = 0; j < 4; ++j) acc1[j] = acc0[j];
                }
  #pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                  float2 w[3][4];
  #pragma unroll
                  for (int dy = 0; dy < 3; ++dy) {
                    const float4 a = *reinterpret_cast<const float4*>(wbase + (dy * 3 + dx) * C::DWS + q0);
                    const float4 e = *reinterpret_cast<const float4*>(wbase + (dy * 3 + dx) * C::DWS + q1);
                    w[dy][0] = make_float2(a.x, a.y); w[dy][1] = make_float2(a.z, a.w);
                    w[dy][2] = make_float2(e.x, e.y); w[dy][3] = make_float2(e.z, e.w);
                  }
  #pragma unroll
                  for (int wr = 0; wr < NR; ++wr) {
                    if (wr >= 3 && !two) continue;                           // rows only the (absent) second pixel needs
                    const float* hp = h0 + (wr * C::HS_COLS + dx) * C::HS_STRIDE;
                    const float4 a = *reinterpret_cast<const float4*>(hp + q0);
                    const float4 e = *reinterpret_cast<const float4*>(hp + q1);
                    const float2 d[4] = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(e.x, e.y), make_float2(e.z, e.w)};
                    if (wr < 3) {
  #pragma unroll
                      for (int j = 0; j < 4; ++j) acc0[j] = ffma2(d[j], w[wr][j], acc0[j]);
                    }
                    if (RPI == 2 && wr >= C::STRIDE) {
  #pragma unroll
                      for (int j = 0; j < 4; ++j) acc1[j] = ffma2(d[j], w[wr - C::STRIDE][j], acc1[j]);
                    }
                  }
                }
                constexpr float kOut = 6.0f * kActScale;                   // relu6(x) * kActScale = sat(x/6) * 384
                const int m2 = f * C::M2F + oy * C::WOP + ox;
                {
                  uint32_t h[4], l[4];
  #pragma unroll
                  for (int j = 0; j < 4; ++j)
                    split2_f16<false>(__saturatef(acc0[j].x) * kOut, __saturatef(acc0[j].y) * kOut, h[j], l[j]);
                  uint8_t* dst = sA2 + (m2 >> 7) * (128 * C::NC * 2) + ((m2 & 127) >> 3) * 128 + kg * 2048 + (m2 & 7) * 16;
                  *reinterpret_cast<uint4*>(dst) = swz ? make_uint4(h[2], h[3], h[0], h[1]) : make_uint4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<uint4*>(dst + C::A2_PLANE) = swz ? make_uint4(l[2], l[3], l[0], l[1]) : make_uint4(l[0], l[1], l[2], l[3]);
                }
                if (two) {
                  const int m3 = m2 + C::WOP;
                  uint32_t h[4], l[4];
  #pragma unroll
                  for (int j = 0; j < 4; ++j)
                    split2_f16<false>(__saturatef(acc1[j].x) * kOut, __saturatef(acc1[j].y) * kOut, h[j], l[j]);
                  uint8_t* dst = sA2 + (m3 >> 7) * (128 * C::NC * 2) + ((m3 & 127) >> 3) * 128 + kg * 2048 + (m3 & 7) * 16;
                  *reinterpret_cast<uint4*>(dst) = swz ? make_uint4(h[2], h[3], h[0], h[1]) : make_uint4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<uint4*>(dst + C::A2_PLANE) = swz ? make_uint4(l[2], l[3], l[0], l[1]) : make_uint4(l[0], l[1], l[2], l[3]);
                }
              }
              it += TPG / 8;
              while (it >= per_kg && kg < kg_end) { it -= per_kg; ++kg; }
            }
          }
        }
        SYN_TRACE(0, c, 5);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&bar_a2));
      }
      SYN_TRACE(0, 63, 1);

      if (tile + (int)gridDim.x < ntiles) prep(tile + gridDim.x);   // XA is free: every GEMM1 of this tile is done
      prefetch_x(tile + 2 * (int)gridDim.x);
      // ---- EPI2: s3*D2 + b3 (+ skip) -> global NHWC --------------------------------------------------
      SYN_TRACE(0, 63, 2);
      constexpr int JW = (C::COUT_P % 32 == 0 && C::MT2 * (C::COUT_P / 32) >= 2 * NWG) ? 32
                         : (C::MT2 * (C::COUT_P / 16) >= 2 * NWG) ? 16 : 8;
      constexpr int JC = C::COUT_P / JW;
      // GEMM2 row m2 -> output pixel of the tile (rows are padded to WOP pixels when DW3 needs it); -1 = no pixel
      auto out_pixel = [&](int m2) -> int {
        if (m2 >= M2) return -1;
        if constexpr (C::WOP == C::WO) return m2;                  // tiles are contiguous in NHWC memory
        const int oyl = m2 / C::WOP, oxl = m2 - oyl * C::WOP;
        return oxl < C::WO ? oyl * C::WO + oxl : -1;
      };
      // The skip connection (stride 1, CIN == COUT: the same pixel of the block input) is a dependent global load per
      // 16 bytes of output: fetch it BEFORE waiting for the last GEMM2, one item ahead of its use afterwards.
      float4 res_cur[JW / 4];
      auto load_res = [&](int e, float4 (&r)[JW / 4]) {
        if constexpr (C::RES) {
          const int t = e / JC, j0 = (e - t * JC) * JW;
          const int pix = out_pixel(t * 128 + row);
#pragma unroll
          for (int j = 0; j < JW; j += 4) {
            r[j / 4] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pix >= 0 && j0 + j < C::COUT)
              r[j / 4] = __ldg(reinterpret_cast<const float4*>(p.x + ((size_t)(f0 * C::W + oy0) * C::W + pix) * C::CIN + j0 + j));
          }
        }
      };
      if constexpr (C::RES) {
        if (wg < mt2 * JC) load_res(wg, res_cur);
      }
      mbar_wait(smem_u32(&bar_g2), n_g2 & 1, p.err);
      ++n_g2;
      tc_fence_after_sync();
      SYN_TRACE(0, 63, 3);
      {
        for (int e = wg; e < mt2 * JC; e += NWG) {
          const int t = e / JC, j0 = (e - t * JC) * JW;
          const int pix = out_pixel(t * 128 + row);
          float* orow = p.y + ((size_t)(f0 * C::WO + oy0) * C::WO + max(pix, 0)) * C::COUT;
          float v[JW];
          const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + C::D2_COL + t * C::COUT_P + j0;
          if constexpr (JW == 32) tmem_ld32(taddr, v); else if constexpr (JW == 16) tmem_ld16(taddr, v); else tmem_ld8(taddr, v);
          if (pix >= 0) {
#pragma unroll
            for (int j = 0; j < JW; j += 4) {
              if (j0 + j < C::COUT) {
                const float4 bb = *reinterpret_cast<const float4*>(sB3 + j0 + j);
                const float4 sc = *reinterpret_cast<const float4*>(sB3 + C::COUT_P + j0 + j);
                float4 o = make_float4(fmaf(v[j], sc.x, bb.x), fmaf(v[j + 1], sc.y, bb.y), fmaf(v[j + 2], sc.z, bb.z),
                                       fmaf(v[j + 3], sc.w, bb.w));
                if constexpr (C::RES) {
                  const float4 r = res_cur[j / 4];
                  o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(orow + j0 + j) = o;
              }
            }
          }
          if constexpr (C::RES) {                                  // next item's skip values: in flight during its TMEM load
            if (e + NWG < mt2 * JC) load_res(e + NWG, res_cur);
          }
        }
      }
      SYN_TRACE(0, 63, 4);
      tc_fence_before_sync();
      mbar_arrive(smem_u32(&bar_d2free));
    }
  } else if (warp == NWW) {
    // =============================== MMA issuer / weight loader ====================================
    // The whole warp runs this control flow convergently and every batch of tcgen05.mma / bulk copies sits
    // under one elect.sync: the compiler then knows a single thread issues them (no per-instruction
    // uniformisation loop, descriptors straight from uniform registers).  Measured with tools/umma_timing:
    // ~170 cycles per MMA from a `tid == X` branch against 32 + N/4 (the operand-read floor) this way.
    const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t total_chunks = (uint32_t)my_tiles * C::NCHUNK;
    auto load_chunk = [&](uint32_t gi) {                     // streaming: chunk gi -> slot gi % WSTAGES
      const uint32_t slot = gi % C::WSTAGES, c = gi % C::NCHUNK;
      mbar_expect_tx(smem_u32(&bar_wfull[slot]), C::CHUNK_BYTES);
      bulk_g2s(smem_u32(sWch + slot * C::CHUNK_BYTES), p.wimg + C::B3_BYTES + (size_t)c * C::CHUNK_BYTES, C::CHUNK_BYTES,
               smem_u32(&bar_wfull[slot]));
    };
    if (elect_one()) {
      if constexpr (C::WSTREAM) {
        mbar_expect_tx(smem_u32(&bar_w), C::B3_BYTES);
        bulk_g2s(smem_u32(smem + C::S_B3), p.wimg, C::B3_BYTES, smem_u32(&bar_w));
        for (uint32_t k = 0; k < (uint32_t)C::WSTAGES; ++k)
          if (total_chunks > k) load_chunk(k);
      } else {
        mbar_expect_tx(smem_u32(&bar_w), C::W_BYTES);
        bulk_g2s(smem_u32(smem + C::S_B3), p.wimg, C::W_BYTES, smem_u32(&bar_w));
      }
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar_w), 0, p.err);
    const uint32_t idesc1 = make_idesc_f16(128, C::NC);
    const uint32_t idesc2 = make_idesc_f16(128, C::N2);
    constexpr uint32_t LBO_W1 = (C::NC / 8) * 128, LBO_W3 = (C::COUT_P / 8) * 128;
    uint32_t n_x = 0, n_epi1 = 0, n_a2 = 0, n_free = 0, n_g2i = 0;
    uint32_t g = 0;                                          // chunk counter of the current GEMM2
    int ntile_local = 0;

    // descriptor halves that never change (SBO = 128 everywhere); per MMA only `lo` moves by (bytes >> 4)
    const uint32_t d_hi = smem_desc_hi(128);
    const uint32_t a2_lo = smem_desc_lo(smem_u32(sA2), 2048);
    const uint32_t w_lo1 = smem_desc_lo(smem_u32(sWch) + C::CH_W1, LBO_W1), w_lo3 = smem_desc_lo(smem_u32(sWch) + C::CH_W3, LBO_W3);
    auto gemm1 = [&](uint32_t gi, int c, int mt1) {
      const int slot = C::WSTREAM ? (int)(gi % C::WSTAGES) : c;
      if constexpr (C::WSTREAM) mbar_wait(smem_u32(&bar_wfull[slot]), (gi / C::WSTAGES) & 1, p.err);
      const uint32_t wb = w_lo1 + ((slot * C::CHUNK_BYTES) >> 4);
      if (elect_one()) {
      for (int t = 0; t < mt1; ++t) {
        const uint32_t xa = tmem + C::XA_COL + t * C::CIN_P;     // A from TMEM: N/2 cycles per MMA, no smem read of X
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          if (pass >= p.npass) break;                              // single-pass engine: hi * hi only
#pragma unroll
          for (int ks = 0; ks < C::CIN_P / 16; ++ks)
            umma_f16_ts(tmem + t * C::NC, xa + (pass == 2 ? C::CIN_P / 2 : 0) + ks * 8,
                        desc64(d_hi, wb + (((pass == 1 ? C::W1_PLANE : 0) + ks * 2 * LBO_W1) >> 4)), idesc1,
                        (pass > 0 || ks > 0) ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(&bar_d1));
      }
      __syncwarp();
    };
    auto gemm2 = [&](uint32_t gi, int c, int mt2) {
      const int slot = C::WSTREAM ? (int)(gi % C::WSTAGES) : c;
      const uint32_t wb = w_lo3 + ((slot * C::CHUNK_BYTES) >> 4);
      const uint32_t acc0 = (c > 0) ? 1u : 0u;
      if (elect_one()) {
      for (int t = 0; t < mt2; ++t) {
        const uint32_t ab = a2_lo + ((t * (128 * C::NC * 2)) >> 4);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          if (pass >= p.npass) break;
#pragma unroll
          for (int ks = 0; ks < C::NC / 16; ++ks)
#pragma unroll
            for (int hh = 0; hh < C::NSPLIT; ++hh)
              umma_f16(tmem + C::D2_COL + t * C::COUT_P + hh * C::N2,
                       desc64(d_hi, ab + (((pass == 2 ? C::A2_PLANE : 0) + ks * 4096) >> 4)),
                       desc64(d_hi, wb + (((pass == 1 ? C::W3_PLANE : 0) + ks * 2 * LBO_W3 + hh * (C::N2 / 8) * 128) >> 4)),
                       idesc2, (pass > 0 || ks > 0) ? 1u : acc0);
        }
      }
      umma_commit(smem_u32(&bar_g2));
      }
      __syncwarp();
    };

    // stem, fp32 crops: bulk-copy (TMA) the crop rows of a tile into sIn, one tile ahead of the workers
    auto stage_rows = [&](int tile) {
      if constexpr (C::STEM) {
        if (p.x_u8 != nullptr || tile >= ntiles) return;
        if (!elect_one()) return;
        const int fgq = tile / C::STRIPS, spq = tile - fgq * C::STRIPS;
        const int iy0q = spq * C::RO * C::STRIDE - 1;
        const int rfq = max(iy0q, 0), rlq = min(iy0q + C::RWIN - 1, C::W - 1);
        const int iy_first = 2 * rfq - 1, nin = 2 * (rlq - rfq + 1) + 1;
        const int r_lo = (iy_first < 0) ? -iy_first : 0;                     // first / last staged row inside the crop
        const int r_hi = min(nin - 1, kImg - 1 - iy_first);
        const uint32_t bytes = (uint32_t)(r_hi - r_lo + 1) * kImg * 4;          // contiguous in the crop and in sIn
        mbar_expect_tx(smem_u32(&bar_in), 3 * bytes);
        for (int ci = 0; ci < 3; ++ci)
          bulk_g2s(smem_u32(sIn + (ci * C::IN_ROWS + r_lo) * C::IN_STRIDE),
                   p.x + ((size_t)(fgq * 3 + ci) * kImg + iy_first + r_lo) * kImg, bytes, smem_u32(&bar_in));
      }
    };
#if SYN_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");     // the stem's crop rows are the previous step's business only in
#endif                                                     // theory (inputs), but the rule is kept uniform
    stage_rows(blockIdx.x);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ntile_local) {
      const int fg = tile / C::STRIPS, sp = tile - fg * C::STRIPS;
      int f0_unused, nfaces;
      group_faces(fg, f0_unused, nfaces);
      const int iy0 = sp * C::RO * C::STRIDE - 1;
      const int rf = max(iy0, 0), rl = min(iy0 + C::RWIN - 1, C::W - 1);
      const int mt1 = (nfaces * (rl - rf + 1) * C::W + 127) >> 7;
      const int mt2 = (nfaces * C::M2F + 127) >> 7;
#ifdef SYN_FUSED_TRACE
      const bool trace_on = blockIdx.x == 0 && tile == (int)gridDim.x && tid == NWT;
#endif
      SYN_TRACE(1, 63, 0);
      mbar_wait(smem_u32(&bar_x), n_x & 1, p.err);
      ++n_x;
      tc_fence_after_sync();
      SYN_TRACE(1, 63, 1);
      gemm1(g, 0, mt1);
      SYN_TRACE(1, 63, 2);
      stage_rows(tile + gridDim.x);          // sIn is free again: the conversion of this tile has consumed it
      for (int c = 0; c < C::NCHUNK; ++c, ++g) {
        SYN_TRACE(1, c, 0);
        if (c + 1 < C::NCHUNK) {
          mbar_wait(smem_u32(&bar_epi1), n_epi1 & 1, p.err);    // D1 drained by the workers
          ++n_epi1;
          tc_fence_after_sync();
          SYN_TRACE(1, c, 1);
          gemm1(g + 1, c + 1, mt1);
          SYN_TRACE(1, c, 2);
        }
        mbar_wait(smem_u32(&bar_a2), n_a2 & 1, p.err);
        ++n_a2;
        if (c == 0 && ntile_local > 0) {                         // D2 of the previous tile drained
          mbar_wait(smem_u32(&bar_d2free), n_free & 1, p.err);
          ++n_free;
        }
        tc_fence_after_sync();
        SYN_TRACE(1, c, 3);
        gemm2(g, c, mt2);
        SYN_TRACE(1, c, 4);
        if constexpr (C::WSTREAM) {
          // the slot of chunk g may be refilled once GEMM2(g) has read W3c (the workers are already past it)
          mbar_wait(smem_u32(&bar_g2), n_g2i & 1, p.err);
          if (g + C::WSTAGES < total_chunks && elect_one()) load_chunk(g + C::WSTAGES);
          __syncwarp();
          SYN_TRACE(1, c, 5);
        }
        ++n_g2i;
      }
      // the last chunk's EPI1 arrival is not consumed above: keep the phase counter in step
      mbar_wait(smem_u32(&bar_epi1), n_epi1 & 1, p.err);
      ++n_epi1;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == NWW) {
    __syncwarp();
    tmem_dealloc<C::TM_COLS>(tmem);
  }
}

// Face groups of a launch: {split, face_groups}.  FACES == 2: full two-face groups fill whole waves of `sms`
// CTAs; the remainder becomes single-face groups when those still fit in one wave.
template <class C>
inline void fused_tile_plan(int batch, int sms, int& split, int& face_groups) {
  if constexpr (C::FACES == 2) {
    const int full = batch / 2, odd = batch & 1, rem = full % sms;
    split = (2 * rem + odd <= sms) ? full - rem : full;
    face_groups = split + (batch - 2 * split);              // every face past the two-face groups is its own group
  } else {
    split = 0;
    face_groups = (batch + C::FACES - 1) / C::FACES;
  }
}

// ---- the instantiations used by the backbone (SURVEY.md section 8(a) shape table) -------------------
//                          CIN CHID NC COUT  W  S  RO FACES RES    STEM   weight ring slots (0 = resident)
using FusedStemB1 = FusedCfg<27, 32, 32, 16, 60, 1, SYN_RO_STEM, 1, false, true, 0>;    // features[0] + features[1]
using FusedB2 = FusedCfg<16, 96, SYN_NC_B2, 24, 60, 2, SYN_RO_B2, 1, false, false, 0>;       // features[2]
using FusedB3 = FusedCfg<24, 144, SYN_NC_B3, 24, 30, 1, SYN_RO_B3, 1, true, false, 0>;      // features[3]
using FusedB4 = FusedCfg<24, 144, SYN_NC_B4, 32, 30, 2, SYN_RO_B4, 1, false, false, 0>;      // features[4]
using FusedB56 = FusedCfg<32, 192, SYN_NC_B56, 32, 15, 1, 15, 1, true, false, 0>;     // features[5], [6]
using FusedB7 = FusedCfg<32, 192, SYN_NC_B7, 64, 15, 2, 8, 1, false, false, 0>;      // features[7]
using FusedB8 = FusedCfg<64, 384, 64, 64, 8, 1, 8, 2, true, false, 3>;         // features[8..10]
using FusedB11 = FusedCfg<64, 384, 64, 96, 8, 1, 8, 2, false, false, 2>;       // features[11]
using FusedB12 = FusedCfg<96, 576, SYN_NC_B12, 96, 8, 1, 8, 2, true, false, SYN_NC_B12 == 32 ? 3 : 2>;        // features[12], [13]
using FusedB14 = FusedCfg<96, 576, SYN_NC_B14, 160, 8, 2, 4, 2, false, false, SYN_NC_B14 == 32 ? 3 : 2>;      // features[14]
// three ring slots for blocks 15/16 (measured -6 %); narrower chunks in deeper rings were slower everywhere else
using FusedB15 = FusedCfg<160, 960, 32, 160, 4, 1, 4, 8, true, false, 3>;      // features[15], [16]
using FusedB17 = FusedCfg<160, 960, SYN_NC_B17, 320, 4, 1, 4, 8, false, false, SYN_NC_B17 == 16 ? 3 : 2>;     // features[17]

}  // namespace syn
