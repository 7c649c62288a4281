// Host-side CTA emulator for the fused tcgen05 kernels (test infrastructure only).
//
// One OS thread per CUDA thread of ONE CTA at a time.  The model is deliberately a
// "maximum latency" machine so that missing synchronisation shows up as wrong results:
//   * cp.async.bulk copies are delivered only when somebody waits on their mbarrier;
//   * tcgen05.mma instructions are queued in issue order and executed only when a wait on an
//     mbarrier that a later tcgen05.commit targets forces them (so overwriting an operand
//     buffer, or reading TMEM, before the covering commit has been waited for reads/produces
//     stale data);
//   * mbarriers follow the PTX phase / pending-count / tx-count rules;
//   * tcgen05.ld checks the warp's TMEM lane-quarter restriction.
// tf32 operands are truncated to 19 bits when read, as the tensor core does.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <chrono>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define WETTS_GLOBAL static inline
#define WETTS_DEVICE static inline
#define WETTS_LAUNCH_BOUNDS(t, b)
#define WETTS_SMEM_DECL(name) uint8_t* name = ::emu::cta()->smem
#define WETTS_TID (::emu::g_tid)
#define WETTS_BID (::emu::cta()->bid)
#define WETTS_NBLK (::emu::cta()->nblk)

#ifndef __CUDACC__
struct alignas(16) float4 {
  float x, y, z, w;
};
struct alignas(16) uint4 {
  uint32_t x, y, z, w;
};
struct alignas(8) float2 {
  float x, y;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#endif

namespace emu {

constexpr uint32_t kSmemBase = 1024;       // dynamic shared memory does not start at address 0 on hardware either
constexpr uint32_t kSmemBytes = 232448;    // 227 KB opt-in limit
constexpr int kMaxBars = 64;

struct MBar {
  uint64_t phase = 0;
  int expected = 0, pending = 0;
  long long tx = 0;
  bool inited = false;
};
struct Op {
  enum Kind { MMA, COMMIT } kind;
  int mma_kind = 32;   // 32: kind::tf32 (K = 8, 4-byte elements), 16: kind::f16 (K = 16, 2-byte elements)
  uint32_t d_tmem = 0, idesc = 0, accumulate = 0, bar = 0;
  uint64_t adesc = 0, bdesc = 0;
};
struct Copy {
  uint32_t dst, bytes, bar;
  const void* src;
};

struct Cta {
  alignas(1024) uint8_t smem[kSmemBytes];
  float tmem[128][512];
  MBar bars[kMaxBars];
  std::mutex m;                 // protects bars, queue, copies
  std::deque<Op> queue;
  std::vector<Copy> copies;
  int nthreads = 0, bid = 0, nblk = 1;
  uint32_t tmem_alloc_cols = 0;
  std::unique_ptr<std::barrier<>> cta_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
  std::atomic<bool> failed{false};
  unsigned long long n_mma = 0;
};

inline Cta*& cta() {
  static Cta* c = nullptr;
  return c;
}
inline thread_local int g_tid = 0;

[[noreturn]] inline void die(const char* msg) {
  fprintf(stderr, "EMU FAILURE (block %d, thread %d): %s\n", cta()->bid, g_tid, msg);
  fflush(stderr);
  _exit(3);
}

inline uint8_t* smem_ptr(uint32_t addr, uint32_t bytes) {
  if (addr < kSmemBase || addr - kSmemBase + bytes > kSmemBytes) die("shared memory access out of range");
  return cta()->smem + (addr - kSmemBase);
}
inline MBar& bar_at(uint32_t addr) {
  if (addr < kSmemBase || (addr & 7) || (addr - kSmemBase) / 8 >= (uint32_t)kMaxBars) die("bad mbarrier address");
  return cta()->bars[(addr - kSmemBase) / 8];
}
// caller holds cta()->m
inline void bar_check_complete(MBar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase += 1;
    b.pending = b.expected;
  }
  if (b.pending < 0) die("mbarrier over-arrived");
}

inline float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

// caller holds cta()->m
inline void exec_mma(const Op& o) {
  Cta* c = cta();
  const uint32_t N = ((o.idesc >> 17) & 0x3F) << 3, M = ((o.idesc >> 24) & 0x1F) << 4;
  if (M != 128) die("only M = 128 is modelled");
  if (N < 16 || N > 256 || (N & 15)) die("idesc: N must be a multiple of 16 in [16, 256] for M = 128");
  const uint32_t afmt = (o.idesc >> 7) & 7, bfmt = (o.idesc >> 10) & 7;
  if (((o.idesc >> 4) & 3) != 1) die("idesc: expected an f32 accumulator");
  const bool f16 = (o.mma_kind == 16);
  if (f16 ? (afmt != 0 || bfmt != 0) : (afmt != 2 || bfmt != 2)) die("idesc: operand formats do not match the MMA kind");
  if ((o.idesc >> 15) & 3) die("idesc: only K-major operands are modelled");
  const uint32_t K = f16 ? 16 : 8, kpc = f16 ? 8 : 4, esz = f16 ? 2 : 4;
  auto field = [](uint64_t d, int sh) { return (uint32_t)((d >> sh) & 0x3FFF) << 4; };
  const uint32_t a0 = field(o.adesc, 0), a_lbo = field(o.adesc, 16), a_sbo = field(o.adesc, 32);
  const uint32_t b0 = field(o.bdesc, 0), b_lbo = field(o.bdesc, 16), b_sbo = field(o.bdesc, 32);
  if (((o.adesc >> 46) & 1) != 1 || ((o.bdesc >> 46) & 1) != 1) die("descriptor version bit missing");
  if ((o.adesc >> 61) != 0 || (o.bdesc >> 61) != 0) die("swizzled layouts are not modelled");
  const uint32_t col0 = o.d_tmem & 0xFFFF;
  if ((o.d_tmem >> 16) != 0) die("MMA accumulator must start at TMEM lane 0");
  if (col0 + N > c->tmem_alloc_cols) die("MMA accumulator outside the TMEM allocation");
  auto elem = [&](uint32_t start, uint32_t lbo, uint32_t sbo, uint32_t row, uint32_t k) {
    const uint32_t addr = start + (k / kpc) * lbo + (row / 8) * sbo + (row % 8) * 16 + (k % kpc) * esz;
    if (f16) {
      _Float16 h;
      memcpy(&h, smem_ptr(addr, 2), 2);
      return (float)h;
    }
    float v;
    memcpy(&v, smem_ptr(addr, 4), 4);
    return tf32_trunc(v);
  };
  std::vector<float> bm((size_t)N * K);
  for (uint32_t n = 0; n < N; ++n)
    for (uint32_t k = 0; k < K; ++k) bm[n * K + k] = elem(b0, b_lbo, b_sbo, n, k);
  for (uint32_t m = 0; m < 128; ++m) {
    float a[16];
    for (uint32_t k = 0; k < K; ++k) a[k] = elem(a0, a_lbo, a_sbo, m, k);
    for (uint32_t n = 0; n < N; ++n) {
      // products of tf32 / f16 values are exact; the tensor core accumulates wide, modelled with double
      double s = 0.0;
      for (uint32_t k = 0; k < K; ++k) s += (double)a[k] * (double)bm[n * K + k];
      float& d = c->tmem[m][col0 + n];
      d = o.accumulate ? (float)((double)d + s) : (float)s;
    }
  }
  c->n_mma += 1;
}

// Deliver everything that a wait on `bar` may legitimately observe: bulk copies that signal it, and
// the MMA queue up to (and including) the first commit that targets it.   caller holds cta()->m
inline void flush_for(uint32_t bar) {
  Cta* c = cta();
  for (size_t i = 0; i < c->copies.size();) {
    if (c->copies[i].bar == bar) {
      const Copy cp = c->copies[i];
      memcpy(smem_ptr(cp.dst, cp.bytes), cp.src, cp.bytes);
      MBar& b = bar_at(bar);
      b.tx -= cp.bytes;
      bar_check_complete(b);
      c->copies.erase(c->copies.begin() + i);
    } else {
      ++i;
    }
  }
  bool has = false;
  for (const Op& o : c->queue) has = has || (o.kind == Op::COMMIT && o.bar == bar);
  while (has) {
    Op o = c->queue.front();
    c->queue.pop_front();
    if (o.kind == Op::MMA) {
      exec_mma(o);
    } else {
      MBar& b = bar_at(o.bar);
      b.pending -= 1;
      bar_check_complete(b);
      if (o.bar == bar) break;
    }
  }
}

// seconds one mbarrier wait may take before it is reported as a deadlock (EMU_TIMEOUT_S; emulated MMAs are slow, and a
// role that waits for a whole work item of another role can legitimately wait long on a loaded machine)
inline int wait_timeout_s() {
  static const int v = getenv("EMU_TIMEOUT_S") ? atoi(getenv("EMU_TIMEOUT_S")) : 20;
  return v;
}
}  // namespace emu

namespace wetts {
namespace tc {

WETTS_DEVICE uint32_t smem_u32(const void* p) {
  return emu::kSmemBase + (uint32_t)((const uint8_t*)p - emu::cta()->smem);
}
WETTS_DEVICE void cta_sync() { emu::cta()->cta_bar->arrive_and_wait(); }
WETTS_DEVICE void warp_sync() { emu::cta()->warp_bars[emu::g_tid >> 5]->arrive_and_wait(); }
WETTS_DEVICE float ldg(const float* p) { return *p; }
WETTS_DEVICE long long clock_now() { return 0; }
WETTS_DEVICE void spin_cycles(long long) {}
WETTS_DEVICE void trap_now() { emu::die("kernel trap"); }
WETTS_DEVICE int ldg_i32(const int* p) { return *p; }
WETTS_DEVICE long long ldg_i64(const long long* p) { return *p; }
WETTS_DEVICE float4 ldg4(const float* p) {
  if ((uintptr_t)p & 15) emu::die("16 B load from a misaligned address");
  return *reinterpret_cast<const float4*>(p);
}

WETTS_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::MBar& b = emu::bar_at(bar);
  b = emu::MBar();
  b.expected = b.pending = (int)count;
  b.inited = true;
}
WETTS_DEVICE void mbar_init_fence() {}
WETTS_DEVICE void mbar_wait(uint32_t bar, uint32_t parity) {
  const auto t0 = std::chrono::steady_clock::now();
  unsigned spins = 0;
  for (;;) {
    {
      std::lock_guard<std::mutex> g(emu::cta()->m);
      emu::MBar& b = emu::bar_at(bar);
      if (!b.inited) emu::die("wait on an uninitialised mbarrier");
      if ((b.phase & 1) != parity) return;
      emu::flush_for(bar);
      if ((b.phase & 1) != parity) return;
    }
    if (++spins > 64) std::this_thread::sleep_for(std::chrono::microseconds(50));
    else std::this_thread::yield();
    if ((spins & 1023) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(emu::wait_timeout_s())) {
      fprintf(stderr, "mbarrier 0x%x parity %u (block %d thread %d)\n", bar, parity, emu::cta()->bid, emu::g_tid);
      if (getenv("EMU_DEADLOCK_DUMP")) std::this_thread::sleep_for(std::chrono::seconds(5));   // let the other waiters report too
      emu::die("mbarrier wait timed out (deadlock)");
    }
  }
}
WETTS_DEVICE void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::MBar& b = emu::bar_at(bar);
  b.tx += bytes;
  b.pending -= 1;
  emu::bar_check_complete(b);
}
WETTS_DEVICE void mbar_arrive(uint32_t bar) {
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::MBar& b = emu::bar_at(bar);
  b.pending -= 1;
  emu::bar_check_complete(b);
}
WETTS_DEVICE void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  if ((dst & 15) || (bytes & 15) || ((uintptr_t)src & 15)) emu::die("cp.async.bulk needs 16 B alignment");
  emu::smem_ptr(dst, bytes);
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::cta()->copies.push_back({dst, bytes, bar, src});
}
WETTS_DEVICE uint64_t l2_policy_evict_last() { return 0; }
WETTS_DEVICE void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t) {
  bulk_g2s(dst, src, bytes, bar);
}
WETTS_DEVICE void st_streaming(float* p, float v) { *p = v; }
WETTS_DEVICE void fence_async_smem() {}
WETTS_DEVICE void tc_fence_before() {}
WETTS_DEVICE void tc_fence_after() {}
WETTS_DEVICE void tc_commit(uint32_t bar) {
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::Op o;
  o.kind = emu::Op::COMMIT;
  o.bar = bar;
  emu::cta()->queue.push_back(o);
}
WETTS_DEVICE bool elect_one() { return (emu::g_tid & 31) == 0; }
WETTS_DEVICE void tc_mma_tf32_x3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                 uint32_t idesc, uint32_t accumulate_first) {
  if ((emu::g_tid & 31) != 0) return;
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::Op o;
  o.kind = emu::Op::MMA;
  o.d_tmem = d_tmem;
  o.idesc = idesc;
  o.adesc = a_lo; o.bdesc = b_hi; o.accumulate = accumulate_first;
  emu::cta()->queue.push_back(o);
  o.adesc = a_hi; o.bdesc = b_lo; o.accumulate = 1;
  emu::cta()->queue.push_back(o);
  o.adesc = a_hi; o.bdesc = b_hi; o.accumulate = 1;
  emu::cta()->queue.push_back(o);
}
WETTS_DEVICE void tc_mma_tf32_split2(uint32_t d_tmem, uint32_t d_tmem_small, uint64_t a_hi, uint64_t a_lo, uint64_t b_hilo,
                                     uint32_t idesc_2n, uint32_t idesc_n, uint32_t accumulate_first) {
  if ((emu::g_tid & 31) != 0) return;
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::Op o;
  o.kind = emu::Op::MMA;
  o.d_tmem = d_tmem; o.idesc = idesc_2n; o.adesc = a_hi; o.bdesc = b_hilo; o.accumulate = accumulate_first;
  emu::cta()->queue.push_back(o);
  o.d_tmem = d_tmem_small; o.idesc = idesc_n; o.adesc = a_lo; o.bdesc = b_hilo; o.accumulate = 1;
  emu::cta()->queue.push_back(o);
}
WETTS_DEVICE void emu_push_mma(int kind, uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  if ((emu::g_tid & 31) != 0) return;
  std::lock_guard<std::mutex> g(emu::cta()->m);
  emu::Op o;
  o.kind = emu::Op::MMA;
  o.mma_kind = kind;
  o.d_tmem = d_tmem; o.idesc = idesc; o.adesc = a; o.bdesc = b; o.accumulate = accumulate;
  emu::cta()->queue.push_back(o);
}
WETTS_DEVICE void tc_mma_tf32_1(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  emu_push_mma(32, d_tmem, a, b, idesc, accumulate);
}
WETTS_DEVICE void tc_mma_f16_1(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  emu_push_mma(16, d_tmem, a, b, idesc, accumulate);
}
WETTS_DEVICE void tc_mma_f16_split2(uint32_t d_tmem, uint32_t d_tmem_small, uint64_t a_hi, uint64_t a_lo, uint64_t b_hilo,
                                    uint32_t idesc_2n, uint32_t idesc_n, uint32_t accumulate_first) {
  emu_push_mma(16, d_tmem, a_hi, b_hilo, idesc_2n, accumulate_first);
  emu_push_mma(16, d_tmem_small, a_lo, b_hilo, idesc_n, 1u);
}
// f16 split pair (see tc_prims.cuh): element 0 in the low half
WETTS_DEVICE uint32_t f16x2_pack(float lo_elem, float hi_elem) {
  auto cv = [](float x) {
    if (x > 65504.f) x = 65504.f;          // satfinite
    if (x < -65504.f) x = -65504.f;
    _Float16 h = (_Float16)x;
    uint16_t u;
    memcpy(&u, &h, 2);
    return (uint32_t)u;
  };
  return cv(lo_elem) | (cv(hi_elem) << 16);
}
WETTS_DEVICE void f16x2_unpack(uint32_t v, float& lo_elem, float& hi_elem) {
  uint16_t a = (uint16_t)(v & 0xFFFF), b = (uint16_t)(v >> 16);
  _Float16 ha, hb;
  memcpy(&ha, &a, 2);
  memcpy(&hb, &b, 2);
  lo_elem = (float)ha;
  hi_elem = (float)hb;
}
WETTS_DEVICE uint32_t warp_uniform(uint32_t v) { return v; }
WETTS_DEVICE uint32_t uniform_bits(uint32_t v, int lo, int hi) { return v & (((hi >= 32) ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u)); }
WETTS_DEVICE void tmem_ld16(uint32_t taddr, float* v) {
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xFFFF;
  const int warp = emu::g_tid >> 5, lane = emu::g_tid & 31;
  if (lane_base != (uint32_t)(32 * (warp & 3))) emu::die("tcgen05.ld: a warp may only read its own TMEM lane quarter");
  if (col + 16 > emu::cta()->tmem_alloc_cols) emu::die("tcgen05.ld outside the TMEM allocation");
  // NOTE: no flush here -- reading accumulators without having waited for the covering commit sees stale data
  std::lock_guard<std::mutex> g(emu::cta()->m);
  for (int i = 0; i < 16; ++i) v[i] = emu::cta()->tmem[lane_base + lane][col + i];
}
WETTS_DEVICE void tmem_ld16_nowait(uint32_t taddr, float* v) { tmem_ld16(taddr, v); }
WETTS_DEVICE void tmem_ld_wait() {}
WETTS_DEVICE void tmem_alloc(uint32_t slot_smem_addr, uint32_t cols) {
  if ((emu::g_tid & 31) != 0) return;
  if (cols < 32 || cols > 512 || (cols & (cols - 1))) emu::die("tcgen05.alloc: columns must be a power of two in [32, 512]");
  emu::cta()->tmem_alloc_cols = cols;
  const uint32_t base = 0;
  memcpy(emu::smem_ptr(slot_smem_addr, 4), &base, 4);
}
WETTS_DEVICE void tmem_dealloc(uint32_t, uint32_t) {}
WETTS_DEVICE float tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if (((u >> 23) & 0xFF) != 0xFF) u += 0x1000u;
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

}  // namespace tc
}  // namespace wetts

namespace emu {

// Runs `kernel(args)` for a grid of `nblk` CTAs of `nthreads` threads, one CTA after the other.
template <typename Kernel, typename Args>
void launch(Kernel kernel, const Args& args, int nblk, int nthreads, unsigned long long* n_mma = nullptr) {
  for (int b = 0; b < nblk; ++b) {
    std::unique_ptr<Cta> c(new Cta());
    memset(c->smem, 0xCD, sizeof(c->smem));       // poison: uninitialised reads become visible
    for (auto& row : c->tmem) for (float& x : row) x = 1.0e30f;
    c->nthreads = nthreads;
    c->bid = b;
    c->nblk = nblk;
    c->cta_bar.reset(new std::barrier<>(nthreads));
    for (int w = 0; w < nthreads / 32; ++w) c->warp_bars.emplace_back(new std::barrier<>(32));
    cta() = c.get();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
      th.emplace_back([&, t]() {
        g_tid = t;
        kernel(args);
      });
    for (auto& t : th) t.join();
    if (!c->queue.empty()) {
      // trailing commits nobody waited for are legal; trailing MMAs without any commit are not
      for (const Op& o : c->queue)
        if (o.kind == Op::MMA) { fprintf(stderr, "EMU: MMAs left in the queue at kernel exit\n"); _exit(3); }
    }
    if (!c->copies.empty()) { fprintf(stderr, "EMU: bulk copies never waited for at kernel exit\n"); _exit(3); }
    if (n_mma) *n_mma += c->n_mma;
    cta() = nullptr;
  }
}

}  // namespace emu
