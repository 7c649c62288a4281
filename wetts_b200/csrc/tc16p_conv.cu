// Launch side of the pipelined per-layer f16 conv kernel (tc16p_conv_kernel.cuh).  It consumes the packed weights and
// the tiling plan of tc16_conv_kernel.cu unchanged.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.cuh"
#include "tc16p_conv_kernel.cuh"

namespace wetts {

static std::atomic<int> g_tc16p{-1};
bool tc16p_enabled() {
  int v = g_tc16p.load();
  if (v < 0) {
    v = getenv("WETTS_TC16P") ? (atoi(getenv("WETTS_TC16P")) != 0) : 0;   // opt-in until measured on hardware
    g_tc16p.store(v);
  }
  return v != 0;
}
void set_tc16p_enabled(bool on) { g_tc16p.store(on ? 1 : 0); }

// true if the launch was taken (large-mode plan whose ring fits); false: the caller uses conv1d_tc16_kernel
bool launch_conv1d_tc16p(const ConvArgs& a, cudaStream_t s) {
  const TcPlan& pl = a.tc16;
  if (pl.mode != 1) return false;
  const int MB = (a.T > 128 && pl.MB == 2) ? 2 : 1;
  // ring depths: weight slots first (up to 4; what the MMAs wait for), then activation slots (4 or 2)
  static const int max_nb = getenv("WETTS_TC16P_NB") ? atoi(getenv("WETTS_TC16P_NB")) : kTc16pNB;
  const size_t cap = 227 * 1024;
  int nb = 2, na = 2;
  if (tc16p_smem_bytes(a.K, a.dil, pl.N, pl.KC, MB, na, nb) > cap) return false;
  while (nb < max_nb && nb < kTc16pNB && tc16p_smem_bytes(a.K, a.dil, pl.N, pl.KC, MB, na, nb + 1) <= cap) ++nb;
  if (tc16p_smem_bytes(a.K, a.dil, pl.N, pl.KC, MB, 4, nb) <= cap) na = 4;
  const size_t smem = tc16p_smem_bytes(a.K, a.dil, pl.N, pl.KC, MB, na, nb);
  TcConvArgs p;
  p.c = a;
  p.wtc = reinterpret_cast<const float*>(a.wtc16);
  const int R = 128 * MB + (a.K - 1) * a.dil;
  p.N = pl.N; p.n_tiles = pl.n_tiles; p.KC = pl.KC; p.n_chunks = pl.n_chunks; p.MB = MB;
  // two accumulator sets when one item (two M blocks) needs at most half of TMEM: ping-pong between MMAs and drain
  // (experiment, off by default: it ran clean on hardware -- 43 parity tests and the benches of profiles/r02i -- but one
  // emulator case with dedicated drain warps deadlocks intermittently, unexplained; the all-warps assignment is the
  // validated one)
  static const int opt_pp = getenv("WETTS_TC16P_PINGPONG") ? atoi(getenv("WETTS_TC16P_PINGPONG")) : 0;
  p.tmem_cols = pl.tmem_cols; p.G = pl.tmem_cols / (MB * 2 * pl.N);
  if (opt_pp && p.G >= 2) { p.G = p.G / 2; p.acc_slots = 2; }
  p.n_abuf = na; p.n_bbuf = nb;
  static const int opt_aw = getenv("WETTS_TC16P_ALLWARPS") ? atoi(getenv("WETTS_TC16P_ALLWARPS")) : 1;
  static const int opt_nt_minor = getenv("WETTS_TC16_NTMINOR") ? atoi(getenv("WETTS_TC16_NTMINOR")) : 1;
  static const int opt_prefetch = getenv("WETTS_TC16_PREFETCH") ? atoi(getenv("WETTS_TC16_PREFETCH")) : 1;
  // 1 (default): all worker warps stage, then drain -- the assignment validated in the emulator.  0: dedicated staging /
  // drain warps (experiment: intermittent deadlocks in the emulator, none seen on hardware).  2: dedicated when two
  // accumulator sets let them overlap the MMAs.
  p.all_warps = (opt_aw == 2) ? (p.acc_slots == 2 ? 0 : 1) : opt_aw;
  p.nt_minor = (opt_nt_minor && pl.n_chunks > 1 && pl.n_tiles > 1) ? 1 : 0;
  p.l2_prefetch = opt_prefetch;
  static const int opt_skip = getenv("WETTS_TC16_DEBUG_SKIP") ? atoi(getenv("WETTS_TC16_DEBUG_SKIP")) : 0;
  p.debug_skip = opt_skip; p.R_pad = (R + 7) & ~7;
  static DynSmemAttr attr, attr_prof;
  const int n_sm = current_device_sm_count();
  if (n_sm <= 0) return true;
  const int group_rows = p.G * 128 * MB;
  const long long items = (long long)a.B * ((a.T + group_rows - 1) / group_rows) * pl.n_tiles;
  const int grid = (int)(items < n_sm ? items : n_sm);
  static const int opt_prof = getenv("WETTS_TC16P_PROFILE") ? atoi(getenv("WETTS_TC16P_PROFILE")) : 0;
  if (opt_prof) {
    // debugging hook: clock64 phase counters of one thread per role; synchronises the stream and prints one line per launch
    if (attr_prof.ensure((const void*)conv1d_tc16p_kernel<true>, smem) != cudaSuccess) return true;
    const size_t n = (size_t)grid * 4 * (kTc16pProfSlots + 1);
    long long* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(long long)) != cudaSuccess) return true;
    cudaMemsetAsync(d, 0, n * sizeof(long long), s);
    p.prof = d;
    conv1d_tc16p_kernel<true><<<grid, kTc16pThreads, smem, s>>>(p);
    count_launch();
    if (cudaStreamSynchronize(s) == cudaSuccess) {
      std::vector<long long> h(n);
      cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
      double m[4][kTc16pProfSlots + 1] = {};
      for (int b = 0; b < grid; ++b)
        for (int r = 0; r < 4; ++r)
          for (int i = 0; i <= kTc16pProfSlots; ++i) m[r][i] += (double)h[((size_t)b * 4 + r) * (kTc16pProfSlots + 1) + i];
      const double per = (double)items / grid * grid;   // items in total
      auto v = [&](int r, int i) { return m[r][i] / per; };
      fprintf(stderr,
              "[tc16p profile] Cin=%d Cout=%d K=%d dil=%d T=%d B=%d ep=%d | N=%d KC=%d chunks=%d MB=%d G=%d slots=%d NA=%d NB=%d aw=%d items/CTA=%.1f | cycles/item: "
              "total=%.0f | MMA: acc_empty=%.0f b_full=%.0f a_full=%.0f issue=%.0f | producer: b_free=%.0f | stager(w2): prefetch=%.0f tile=%.0f (a_free=%.0f) "
              "fence+arrive=%.0f | drain(w15): acc_full=%.0f body=%.0f\n",
              a.Cin, a.Cout, a.K, a.dil, a.T, a.B, (int)a.ep.mode, p.N, p.KC, p.n_chunks, p.MB, p.G, p.acc_slots, na, nb, p.all_warps,
              (double)items / grid, v(0, kTc16pProfSlots), v(0, 0), v(0, 1), v(0, 2), v(0, 3), v(1, 0), v(2, 5), v(2, 2), v(2, 1), v(2, 3),
              v(3, 6), v(3, 7));
    }
    cudaFree(d);
    return true;
  }
  if (attr.ensure((const void*)conv1d_tc16p_kernel<false>, smem) != cudaSuccess) return true;   // error recorded; nothing launched
  conv1d_tc16p_kernel<false><<<grid, kTc16pThreads, smem, s>>>(p);
  count_launch();
  return true;
}

int tc16p_install_fault_word(unsigned int* word) { return tc::install_fault_word_tu(word) == cudaSuccess ? 0 : 1; }

}  // namespace wetts
