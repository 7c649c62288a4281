// Launch side of the row-block-resident per-layer f16 conv kernel (tc16r_conv_kernel.cuh).  It consumes the packed
// weights and the plan of the two-CTA tiling of tc16_conv_kernel.cu (N <= 128, one M block, same KC / chunk count), so a
// layer planned that way takes either route at launch time.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kernels.cuh"
#include "tc16r_conv_kernel.cuh"

namespace wetts {

static std::atomic<int> g_tc16r{-1};
bool tc16r_enabled() {
  int v = g_tc16r.load();
  if (v < 0) {
    // default since it was measured (profiles/r02u_*): 71.6 -> 69.2 ms per step, every multi-tile layer type faster
    v = getenv("WETTS_TC16R") ? (atoi(getenv("WETTS_TC16R")) != 0) : 1;
    g_tc16r.store(v);
  }
  return v != 0;
}

// true if the launch was taken; false: the caller uses conv1d_tc16_kernel
bool launch_conv1d_tc16r(const ConvArgs& a, cudaStream_t s) {
  const TcPlan& pl = a.tc16;
  static const int min_tiles = getenv("WETTS_TC16R_MIN_TILES") ? atoi(getenv("WETTS_TC16R_MIN_TILES")) : 2;
  if (pl.mode != 2 || pl.MB != 1 || pl.n_tiles < min_tiles || pl.n_tiles * pl.N > 1024 || a.T < 128) return false;
  const size_t cap = 227 * 1024;
  int nb = 2;
  if (tc16r_smem_bytes(a.K, a.dil, pl.N, pl.KC, pl.n_chunks, nb) > cap) return false;
  while (nb < kTc16rNB && tc16r_smem_bytes(a.K, a.dil, pl.N, pl.KC, pl.n_chunks, nb + 1) <= cap) ++nb;
  const size_t smem = tc16r_smem_bytes(a.K, a.dil, pl.N, pl.KC, pl.n_chunks, nb);
  TcConvArgs p;
  p.c = a;
  p.wtc = reinterpret_cast<const float*>(a.wtc16);
  const int R = 128 + (a.K - 1) * a.dil;
  p.N = pl.N; p.n_tiles = pl.n_tiles; p.KC = pl.KC; p.n_chunks = pl.n_chunks; p.MB = 1; p.G = 1;
  p.tmem_cols = 512; p.n_abuf = 0; p.n_bbuf = nb; p.R_pad = (R + 7) & ~7;
  static const int opt_prefetch = getenv("WETTS_TC16_PREFETCH") ? atoi(getenv("WETTS_TC16_PREFETCH")) : 1;
  p.l2_prefetch = opt_prefetch;
  const int n_sm = current_device_sm_count();
  if (n_sm <= 0) return true;
  const long long items = (long long)a.B * ((a.T + 127) / 128);
  const int grid = (int)(items < n_sm ? items : n_sm);
  static const int opt_prof = getenv("WETTS_TC16R_PROFILE") ? atoi(getenv("WETTS_TC16R_PROFILE")) : 0;
  if (opt_prof && (a.ep.mode == EPI_GATE || a.ep.mode == EPI_RES_SKIP || a.ep.mode == EPI_PLAIN || a.ep.mode == EPI_CONVT)) {
    // debugging hook: clock64 phase counters of one thread per role; synchronises the stream, one line per launch
    static DynSmemAttr attr_prof[7];
    const size_t n = (size_t)grid * 3 * (kTc16rProfSlots + 1);
    long long* d = nullptr;
    if (cudaMalloc(&d, n * sizeof(long long)) != cudaSuccess) return true;
    cudaMemsetAsync(d, 0, n * sizeof(long long), s);
    p.prof = d;
    cudaError_t e = cudaSuccess;
    if (a.ep.mode == EPI_GATE) { e = attr_prof[3].ensure((const void*)conv1d_tc16r_kernel<EPI_GATE, true>, smem); if (e == cudaSuccess) conv1d_tc16r_kernel<EPI_GATE, true><<<grid, kTc16rThreads, smem, s>>>(p); }
    else if (a.ep.mode == EPI_RES_SKIP) { e = attr_prof[4].ensure((const void*)conv1d_tc16r_kernel<EPI_RES_SKIP, true>, smem); if (e == cudaSuccess) conv1d_tc16r_kernel<EPI_RES_SKIP, true><<<grid, kTc16rThreads, smem, s>>>(p); }
    else if (a.ep.mode == EPI_PLAIN) { e = attr_prof[0].ensure((const void*)conv1d_tc16r_kernel<EPI_PLAIN, true>, smem); if (e == cudaSuccess) conv1d_tc16r_kernel<EPI_PLAIN, true><<<grid, kTc16rThreads, smem, s>>>(p); }
    else { e = attr_prof[6].ensure((const void*)conv1d_tc16r_kernel<EPI_CONVT, true>, smem); if (e == cudaSuccess) conv1d_tc16r_kernel<EPI_CONVT, true><<<grid, kTc16rThreads, smem, s>>>(p); }
    count_launch();
    if (e == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess) {
      std::vector<long long> h(n);
      cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
      double m[3][kTc16rProfSlots + 1] = {};
      for (int b = 0; b < grid; ++b)
        for (int r = 0; r < 3; ++r)
          for (int i = 0; i <= kTc16rProfSlots; ++i) m[r][i] += (double)h[((size_t)b * 3 + r) * (kTc16rProfSlots + 1) + i];
      const double per = (double)items;
      auto v = [&](int r, int i) { return m[r][i] / per; };
      fprintf(stderr,
              "[tc16r profile] Cin=%d Cout=%d K=%d dil=%d T=%d B=%d ep=%d | N=%d tiles=%d KC=%d chunks=%d NB=%d items/CTA=%.1f | cycles per 128-row item: total=%.0f | "
              "MMA: a_ready=%.0f acc_empty=%.0f b_full=%.0f issue=%.0f | drain warp 4: stage=%.0f acc_full=%.0f body=%.0f | drain warp 15: stage=%.0f acc_full=%.0f body=%.0f\n",
              a.Cin, a.Cout, a.K, a.dil, a.T, a.B, (int)a.ep.mode, p.N, p.n_tiles, p.KC, p.n_chunks, nb, (double)items / grid, v(0, kTc16rProfSlots),
              v(0, 0), v(0, 1), v(0, 2), v(0, 3), v(1, 4), v(1, 5), v(1, 6), v(2, 4), v(2, 5), v(2, 6));
    }
    cudaFree(d);
    return true;
  }
  static DynSmemAttr attr[7];
  bool ok = false;
#define WETTS_TC16R_LAUNCH(M)                                                                              \
  case M:                                                                                                  \
    if (attr[M].ensure((const void*)conv1d_tc16r_kernel<M>, smem) != cudaSuccess) return true;             \
    conv1d_tc16r_kernel<M><<<grid, kTc16rThreads, smem, s>>>(p);                                           \
    ok = true;                                                                                             \
    break;
  switch (a.ep.mode) {
    WETTS_TC16R_LAUNCH(EPI_PLAIN)
    WETTS_TC16R_LAUNCH(EPI_RESID)
    WETTS_TC16R_LAUNCH(EPI_MRF)
    WETTS_TC16R_LAUNCH(EPI_GATE)
    WETTS_TC16R_LAUNCH(EPI_RES_SKIP)
    WETTS_TC16R_LAUNCH(EPI_COUPLING)
    WETTS_TC16R_LAUNCH(EPI_CONVT)
    default:
      break;
  }
#undef WETTS_TC16R_LAUNCH
  if (!ok) return false;
  count_launch();
  return true;
}

int tc16r_install_fault_word(unsigned int* word) { return tc::install_fault_word_tu(word) == cudaSuccess ? 0 : 1; }

}  // namespace wetts
