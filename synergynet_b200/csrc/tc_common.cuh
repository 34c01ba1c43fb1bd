// tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a, as inline PTX.
//
// Operand layout used throughout: K-major, SWIZZLE_NONE ("interleaved") canonical UMMA layout.
// A tile of R rows x K columns of bf16 is stored as 8x8 core matrices of 128 contiguous bytes
// (8 rows x 16 bytes); element (r, k) lives at byte
//     (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
// with LBO = stride between core matrices along K and SBO = stride between 8-row groups
// (cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE: ((8,n),2):((1,SBO),LBO) in
// uint128 units).  One tcgen05.mma consumes K = 16 (two core matrices along K).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace syn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) --------------------------------
// bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout (0 = none)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// The two 32-bit halves separately: the MMA issuer keeps `hi` and a base `lo` in registers and only
// adds (byte offset >> 4) to `lo` per instruction instead of rebuilding the 64-bit descriptor.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ uint64_t desc64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// ---- instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16, 16-bit x 16-bit -> f32 -------
__host__ __device__ constexpr uint32_t make_idesc_16b(int M, int N, uint32_t ab_format) {
  return (1u << 4)                     // c_format  = F32
         | (ab_format << 7)            // a_format  (0 = F16, 1 = BF16)
         | (ab_format << 10)           // b_format
         | (0u << 15) | (0u << 16)     // a_major = b_major = K
         | ((uint32_t)(N >> 3) << 17)  // n_dim
         | ((uint32_t)(M >> 4) << 24); // m_dim
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) { return make_idesc_16b(M, N, 1u); }
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) { return make_idesc_16b(M, N, 0u); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU (a hung box is a strike).  After ~2 s of wall
// clock (or as soon as any other thread has already timed out) the caller-provided sticky flag is
// raised and the wait returns, so the kernel drains with garbage instead of spinning forever.
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Slow path: plain try_wait polling (a suspend-time hint was tried and measured ~6 % slower end to end:
// wake-up latency matters more than the issue slots the polling warps take).  The sticky flag (a global
// load) and the wall clock are only looked at every 256 polls.
__device__ __forceinline__ bool mbar_wait_spin(uint32_t bar, uint32_t parity, int* err_flag) {
  const uint64_t t0 = globaltimer_ns();
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((it & 255u) == 0) {
      if (err_flag != nullptr && *reinterpret_cast<volatile int*>(err_flag) != 0) return false;
      if (globaltimer_ns() - t0 > 2000000000ull) {
        // the flag lives in mapped pinned host memory (the host reads it without a device sync): plain store
        if (err_flag != nullptr) { *reinterpret_cast<volatile int*>(err_flag) = 1; __threadfence_system(); }
        return false;
      }
    }
  }
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, int* err_flag) {
  return mbar_wait_spin(bar, parity, err_flag);
}
// SYN_MBAR_INLINE: a real call in a kernel makes ptxas keep the global-memory descriptor in a vector register and copy it
// to a uniform register pair (2 x R2UR) in front of every LDG / STG; the spin loop inlined costs less code than that.
#ifndef SYN_MBAR_INLINE
#define SYN_MBAR_INLINE 0
#endif
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err_flag) {
  if (mbar_try_wait(bar, parity)) return true;
#if SYN_MBAR_INLINE
  return mbar_wait_spin(bar, parity, err_flag);
#else
  return mbar_wait_slow(bar, parity, err_flag);
#endif
}
// always-inline flavour for kernels whose hot loop is made of global stores (dense_recon_fm_kernel)
__device__ __forceinline__ bool mbar_wait_inl(uint32_t bar, uint32_t parity, int* err_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  return mbar_wait_spin(bar, parity, err_flag);
}

// ---- proxies / fences ----------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (tcgen05.mma / bulk copies)
// One lane of a converged warp (elect.sync, full mask).  Code under `if (elect_one())` is known to the
// compiler to run on a single thread, which is what lets tcgen05.mma take uniform-register operands directly.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- TMEM allocation (one full warp executes these) ----------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr) {
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "TMEM cols");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ---- MMA issue / commit (ONE thread) --------------------------------------------------------------
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (lane = row m, each 32-bit column holds the K pair 2j | 2j+1 << 16; one MMA consumes
// 8 columns = K 16), B from shared memory: no shared-memory read of the 4 KB A tile per MMA.
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- TMEM -> registers: warp w reads lanes 32*(w%4)..+31, 16 consecutive fp32 columns ---------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// three 16-column loads in flight, one wait (the dense epilogue reads x, y, z accumulators of the same lane)
__device__ __forceinline__ void tmem_ld16x3(uint32_t t0, uint32_t t1, uint32_t t2, float (&a)[16], float (&b)[16], float (&c)[16]) {
  uint32_t r[3][16];
  const uint32_t ta[3] = {t0, t1, t2};
#pragma unroll
  for (int q = 0; q < 3; ++q)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[q][0]), "=r"(r[q][1]), "=r"(r[q][2]), "=r"(r[q][3]), "=r"(r[q][4]), "=r"(r[q][5]), "=r"(r[q][6]), "=r"(r[q][7]),
          "=r"(r[q][8]), "=r"(r[q][9]), "=r"(r[q][10]), "=r"(r[q][11]), "=r"(r[q][12]), "=r"(r[q][13]), "=r"(r[q][14]), "=r"(r[q][15])
        : "r"(ta[q])
        : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[0][i]); b[i] = __uint_as_float(r[1][i]); c[i] = __uint_as_float(r[2][i]); }
}
// registers -> TMEM: thread t of warp w writes lane 32*(w%4)+t, 4 consecutive columns
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 8 columns without the wait: issue several, then tmem_wait_ld() once (hides the TMEM read latency)
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 consecutive columns with one instruction + one wait
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map), completes on an mbarrier ------
// size multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Same copy with an L2 eviction-priority hint.  evict_last keeps a re-read working set (the 40 MB dense basis image,
// read once per 64-face tile) resident in the 126 MB L2 while a write-once stream several times its size flows through.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}

// ---- packed fp32 arithmetic (sm_100: FFMA2 / FMUL2, two IEEE fp32 lanes per instruction) -----------------
// Bit-identical to two scalar fmaf / multiplies; halves the FMA-pipe issue slots of the depthwise phase.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// ---- fp32 -> (hi, lo) bf16 split: x ~= hi + lo with |x - hi - lo| <= 2^-17 |x| ----------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// ---- fp32 -> (hi, lo) fp16 split: 11 + 11 mantissa bits, |x - hi - lo| <= 2^-22 |x| while lo stays a
// normal fp16 number.  Callers pre-scale by a power of two (kActScale for activations, a per-channel
// scale for weights) so that this holds over the value range that matters, and clamp to the fp16 range.
constexpr float kActScale = 64.0f;           // relu6 range [0,6] -> [0,384]; block inputs: |x| < 1000
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -60000.f), 60000.f);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_f16x2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
// split two fp32 values (already multiplied by their power-of-two scale) into packed hi and lo words;
// CLAMP = false when the inputs are known to be inside the fp16 range (ReLU6 outputs x kActScale)
template <bool CLAMP = true>
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (CLAMP) {
    a = fminf(fmaxf(a, -60000.f), 60000.f);
    b = fminf(fmaxf(b, -60000.f), 60000.f);
  }
  const __half2 h = __floats2half2_rn(a, b);                 // one cvt.rn.f16x2.f32
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

}  // namespace tc
}  // namespace syn
