// tcgen05 / TMA / mbarrier primitives used by the fused kernels (sm_100a inline PTX).
//
// The same kernel source also compiles for the host-side CTA emulator (tests/emu/): with
// WETTS_EMULATE defined every primitive below maps to a functional model in emu_runtime.h
// (one OS thread per CUDA thread, mbarriers with phase/tx semantics, tcgen05.mma evaluated
// from the shared-memory descriptors).  The emulator checks index arithmetic, descriptor
// construction, barrier phases and buffer hand-offs on the CPU; it is test infrastructure
// only and never part of libwetts_b200.so.
#pragma once
#include <stdint.h>

#ifdef WETTS_EMULATE
#include "emu_runtime.h"
#else
#include <cuda_runtime.h>
#include <cstdio>

#define WETTS_GLOBAL __global__
#define WETTS_DEVICE __device__ __forceinline__
#define WETTS_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define WETTS_SMEM_DECL(name) extern __shared__ __align__(128) uint8_t name[]
#define WETTS_TID ((int)threadIdx.x)
#define WETTS_BID ((int)blockIdx.x)
#define WETTS_NBLK ((int)gridDim.x)

namespace wetts {
namespace tc {

WETTS_DEVICE uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
WETTS_DEVICE void cta_sync() { __syncthreads(); }
WETTS_DEVICE void warp_sync() { __syncwarp(); }
WETTS_DEVICE float ldg(const float* p) { return __ldg(p); }
WETTS_DEVICE long long clock_now() { return clock64(); }
WETTS_DEVICE void spin_cycles(long long n) {     // busy-wait n SM cycles (start-up stagger of co-resident CTAs)
  const long long t0 = clock64();
  while (clock64() - t0 < n) {}
}
WETTS_DEVICE void trap_now() { __trap(); }
WETTS_DEVICE float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
WETTS_DEVICE int ldg_i32(const int* p) { return __ldg(p); }
WETTS_DEVICE long long ldg_i64(const long long* p) { return __ldg(p); }

WETTS_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
WETTS_DEVICE void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Watchdog.  A pipeline bug must fail fast instead of hanging the GPU: after 2^24 failed polls the wait records the
// reason in a host-visible word (mapped pinned memory, one per process, installed per translation unit and device by
// install_fault_word_tu(); it stays readable after the context is lost) and traps -> launch failure; the host reports
// "pipeline watchdog fired" instead of an anonymous launch failure (engine.cu take_fault()).
// A non-trapping variant (record the fault, fall through, let the kernel drain) was built and measured in SASS: it
// costs the issue path its uniformity -- the wait loop then has an exit that is not guarded by the try_wait predicate,
// ptxas treats everything after the wait as potentially divergent, every later tcgen05.mma is predicated and its
// operands go through R2UR (fused_resblock2: 27 -> 449 R2UR, per-layer kernel: 16 -> 231).  The dead-end form below
// keeps the control-flow shape of a plain bounded spin.  The loop lives inside ONE asm statement for the same reason
// (a C++ loop around try_wait: 240 -> 57 R2UR in the fused kernel, DESIGN.md "issue path").
__constant__ unsigned int* c_fault_word = nullptr;   // per translation unit and device
static inline cudaError_t install_fault_word_tu(unsigned int* word) {   // call once per device from each .cu that waits
  return cudaMemcpyToSymbol(c_fault_word, &word, sizeof(word));
}
WETTS_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned int* fw = c_fault_word;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 n, f;\n\t"
      "mov.u32 n, 0;\n"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.le.u32 q, n, 0x1000000;\n\t"
      "@q bra WAIT_LOOP;\n\t"
      "setp.ne.u64 q, %2, 0;\n\t"
      "mov.u32 f, 1;\n\t"
      "@q st.volatile.global.u32 [%2], f;\n\t"        // tell the host why
      "trap;\n"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity), "l"(fw)
      : "memory");
}
WETTS_DEVICE void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
WETTS_DEVICE void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
WETTS_DEVICE void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 cache policies: the weight stream is re-read by every CTA for every work item and must survive the
// multi-GB activation stream that flows through the same L2 (evict_last); outputs are written once (stcs).
WETTS_DEVICE uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
WETTS_DEVICE void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
WETTS_DEVICE void st_streaming(float* p, float v) { __stcs(p, v); }
WETTS_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
WETTS_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
WETTS_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
WETTS_DEVICE void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// true in exactly one (converged) lane of the warp; the surrounding control flow stays warp-uniform
WETTS_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// The three 3xTF32 MMAs of one k-step (lo*hi, hi*lo, hi*hi: small terms first), issued by one elected
// lane without a C++ branch so the caller's loop is straight-line warp-uniform code.
WETTS_DEVICE void tc_mma_tf32_x3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                 uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %6, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %3, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %4, %5, 1;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %3, %5, 1;\n\t}" ::"r"(d_tmem),
      "l"(a_hi), "l"(a_lo), "l"(b_hi), "l"(b_lo), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// 3xTF32 with the weight operand stored as [hi | lo] along N (2N rows): two MMAs per k-step instead of three.
//   D[:, 0:2N]  (+)= A_hi * [B_hi | B_lo]      (columns 0..N-1: hi*hi, columns N..2N-1: hi*lo)
//   D[:, N:2N]   += A_lo * B_hi                (small terms accumulate together)
// The activation tile (128 rows x 32 B, the operand that bounds SS-mode MMAs at 128 B/clk of shared-memory
// bandwidth) is read twice instead of three times.  The epilogue adds the two column halves.
WETTS_DEVICE void tc_mma_tf32_split2(uint32_t d_tmem, uint32_t d_tmem_small, uint64_t a_hi, uint64_t a_lo, uint64_t b_hilo,
                                     uint32_t idesc_2n, uint32_t idesc_n, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %7, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %4, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%1], %3, %4, %6, 1;\n\t}" ::"r"(d_tmem),
      "r"(d_tmem_small), "l"(a_hi), "l"(a_lo), "l"(b_hilo), "r"(idesc_2n), "r"(idesc_n), "r"(accumulate_first)
      : "memory");
}
// one MMA of either kind, issued by one elected lane (used by tools/ubench and the f16-split kernels)
WETTS_DEVICE void tc_mma_tf32_1(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pa;\n\t}" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
WETTS_DEVICE void tc_mma_f16_1(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %4, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// f16 form of tc_mma_tf32_split2 (K = 16 channels per MMA): D[:, 0:2N] (+)= A_hi * [B_hi | B_lo'],  D[:, N:2N] += A_lo' * B_hi
WETTS_DEVICE void tc_mma_f16_split2(uint32_t d_tmem, uint32_t d_tmem_small, uint64_t a_hi, uint64_t a_lo, uint64_t b_hilo,
                                    uint32_t idesc_2n, uint32_t idesc_n, uint32_t accumulate_first) {
  asm volatile(
      "{\n\t.reg .pred pe, pa;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pa, %7, 0;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %4, %5, pa;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %6, 1;\n\t}" ::"r"(d_tmem),
      "r"(d_tmem_small), "l"(a_hi), "l"(a_lo), "l"(b_hilo), "r"(idesc_2n), "r"(idesc_n), "r"(accumulate_first)
      : "memory");
}
// A value every lane of the warp holds anyway, routed through a shuffle so that the compiler can PROVE it
// warp-uniform: descriptor arithmetic then stays in uniform registers (UIADD3/UMOV feeding UTCHMMA directly).
// Without this every tcgen05.mma operand takes an R2UR round trip (~90 cycles per MMA measured, see
// tools/ubench/mma_ubench.cu) and small-N MMAs become issue-bound.
WETTS_DEVICE uint32_t warp_uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// Stronger form for small values: rebuild bits [lo, hi) from warp votes.  A vote result is a uniform predicate
// (VOTEU -> UP), so the value is born in the uniform datapath; a shuffle result lives in a vector register and
// ptxas may still convert it with one predicated R2UR per use.
WETTS_DEVICE uint32_t uniform_bits(uint32_t v, int lo, int hi) {
  uint32_t r = 0;
#pragma unroll
  for (int b = lo; b < hi; ++b)
    if (__any_sync(0xffffffffu, (v >> b) & 1u)) r |= (1u << b);
  return r;
}
// 32 lanes x 16 consecutive fp32 columns of TMEM (the warp's own lane quarter)
WETTS_DEVICE void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// same load without the wait: issue several, then tmem_ld_wait() once
WETTS_DEVICE void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
WETTS_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// warp-collective (call from one full warp)
WETTS_DEVICE void tmem_alloc(uint32_t slot_smem_addr, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
WETTS_DEVICE void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
WETTS_DEVICE float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// two fp32 -> packed fp16 pair (round to nearest even, saturating to +-65504 instead of inf); element 0 in the low half
WETTS_DEVICE uint32_t f16x2_pack(float lo_elem, float hi_elem) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi_elem), "f"(lo_elem));
  return d;
}
WETTS_DEVICE void f16x2_unpack(uint32_t v, float& lo_elem, float& hi_elem) {
  asm("{\n\t.reg .b16 l, h;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "cvt.f32.f16 %0, l;\n\t"
      "cvt.f32.f16 %1, h;\n\t}"
      : "=f"(lo_elem), "=f"(hi_elem)
      : "r"(v));
}

}  // namespace tc
}  // namespace wetts
#endif  // WETTS_EMULATE

namespace wetts {
namespace tc {

// Warm L2 with the 128-byte lines that hold floats [t0, t1) of `nrows` rows spaced `row_stride` floats apart (fire and
// forget; the caller's threads share the lines round-robin).  Used one work item ahead by the per-layer kernels: the
// staging loads of an item then pay an L2 hit, not a DRAM round trip, per chunk.
WETTS_DEVICE void l2_prefetch_rows(const float* base, long long row_stride, int nrows, int t0, int t1, int tid, int nthreads) {
#ifndef WETTS_EMULATE
  if (t1 <= t0 || nrows <= 0) return;
  const int lines = ((t1 - t0) >> 5) + 2;                  // an unaligned run of n floats touches at most n/32 + 2 lines
  const int total = nrows * lines;
  for (int i = tid; i < total; i += nthreads) {
    const int r = i / lines, l = i - r * lines;
    const float* row = base + (long long)r * row_stride;
    const uintptr_t line = (reinterpret_cast<uintptr_t>(row + t0) & ~(uintptr_t)127) + (uintptr_t)l * 128u;
    if (line < reinterpret_cast<uintptr_t>(row + t1)) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
  }
#endif
}

// fp32 -> two fp16 operands with the same 22 significand bits as the 3xTF32 split:  x ~ hi + lo' * 2^-11,
// hi = f16(x), lo' = f16((x - hi) * 2^11).  The lo' parts of BOTH operands carry the 2^11 scale, so the two small
// products (hi*lo', lo'*hi) accumulate in their own accumulator columns and are scaled by 2^-11 once, in the
// epilogue.  |x| must stay below 65504 (saturating conversion; activations of this model family are O(10)).
constexpr float kF16LoScale = 2048.0f, kF16LoInv = 1.0f / 2048.0f;
WETTS_DEVICE void f16_split2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  hi2 = f16x2_pack(a, b);
  float ha, hb;
  f16x2_unpack(hi2, ha, hb);
  lo2 = f16x2_pack((a - ha) * kF16LoScale, (b - hb) * kF16LoScale);
}
WETTS_DEVICE void f16_join2(uint32_t hi2, uint32_t lo2, float& a, float& b) {
  float ha, hb, la, lb;
  f16x2_unpack(hi2, ha, hb);
  f16x2_unpack(lo2, la, lb);
  a = ha + la * kF16LoInv;
  b = hb + lb * kF16LoInv;
}

// shared-memory matrix descriptor: K-major, no swizzle (UMMA SmemDescriptor, version 1).
// Element (row r, k) of the operand lives at start + (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4.
WETTS_DEVICE uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
WETTS_DEVICE uint64_t desc_with_lo(uint64_t base, uint32_t lo) {
  return (base & 0xFFFFFFFF00000000ull) | (uint64_t)lo;
}
// kind::tf32 instruction descriptor: fp32 accumulate, tf32 A/B (both K-major), M = 128, N
WETTS_DEVICE uint32_t idesc_tf32_m128(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}
// kind::f16 instruction descriptor: fp32 accumulate, fp16 A/B (format 0, both K-major), M = 128, N; K = 16 per MMA
WETTS_DEVICE uint32_t idesc_f16_m128(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

}  // namespace tc
}  // namespace wetts
