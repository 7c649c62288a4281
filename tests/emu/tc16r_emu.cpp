// CPU check of conv1d_tc16p_kernel (wetts_b200/csrc/tc16r_conv_kernel.cuh) in the CTA emulator against an fp64 Conv1d:
// multi-chunk K loop, taps and dilation, two M blocks, several N tiles, ragged T, input / output masks, leaky-relu
// pre-activation, the plain / residual / gate epilogues, and zero-tile items of the length-aware mode.
//   usage: tc16p_emu Cin Cout K dil B T N KC grid mode(0 plain+relu, 1 resid, 2 gate) [length-aware 0|1]
#define WETTS_EMULATE 1
#include <math.h>

#include <random>

#include "tc16r_conv_kernel.cuh"

using namespace wetts;

int main(int argc, char** argv) {
  if (argc < 11) { printf("usage: %s Cin Cout K dil B T N KC grid mode [la]\n", argv[0]); return 64; }
  const int Cin = atoi(argv[1]), Cout = atoi(argv[2]), K = atoi(argv[3]), dil = atoi(argv[4]), B = atoi(argv[5]), T = atoi(argv[6]);
  const int N = atoi(argv[7]), KC = atoi(argv[8]), grid = atoi(argv[9]), mode = atoi(argv[10]);
  const int la = argc > 11 ? atoi(argv[11]) : 0;
  const int cin16 = (Cin + 15) / 16 * 16, n_chunks = (cin16 + KC - 1) / KC, n_tiles = (Cout + N - 1) / N;
  if (N > 128 || (N % 32) || (KC % 16)) { printf("bad tiling\n"); return 64; }
  std::mt19937 rng(7 + Cin + Cout + T);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> x((size_t)B * Cin * T), w((size_t)Cout * Cin * K), bias(Cout), resid((size_t)B * Cout * T), out((size_t)B * Cout * T, -555.f);
  for (auto& v : x) v = nd(rng);
  for (auto& v : w) v = nd(rng) / sqrtf((float)(Cin * K));
  for (auto& v : bias) v = nd(rng) * 0.1f;
  for (auto& v : resid) v = nd(rng);
  std::vector<long long> len(B), la_len(B);
  for (int b = 0; b < B; ++b) { len[b] = T - 37 * b > 1 ? T - 37 * b : 1; la_len[b] = (b & 1) ? 40 : T; }
  // packed weights: [nt][chunk][tap][kg = KC/8][hl: 2N rows][8]   (pack_conv_tc16_kernel)
  const size_t packed_halfs = (size_t)n_tiles * n_chunks * K * (KC / 8) * 2 * N * 8;
  uint16_t* packed = (uint16_t*)aligned_alloc(128, packed_halfs * 2 + 128);
  for (size_t i = 0; i < packed_halfs; ++i) {
    size_t r = i;
    const int e = (int)(r % 8); r /= 8;
    const int n = (int)(r % N); r /= N;
    const int hl = (int)(r % 2); r /= 2;
    const int kg = (int)(r % (KC / 8)); r /= (KC / 8);
    const int tap = (int)(r % K); r /= K;
    const int chunk = (int)(r % n_chunks); r /= n_chunks;
    const int nt = (int)r;
    const int co = nt * N + n, ci = chunk * KC + kg * 8 + e;
    const float v = (co < Cout && ci < Cin) ? w[((size_t)co * Cin + ci) * K + tap] : 0.f;
    uint32_t hi2, lo2;
    tc::f16_split2(v, 0.f, hi2, lo2);
    packed[i] = (uint16_t)((hl ? lo2 : hi2) & 0xFFFF);
  }
  TcConvArgs p;
  ConvArgs& a = p.c;
  a.in = x.data(); a.in_bs = (long long)Cin * T; a.in_cs = T; a.bias = bias.data();
  a.B = B; a.Cin = Cin; a.Cout = Cout; a.CoutPad = Cout; a.T = T; a.K = K; a.dil = dil; a.pad_left = (K - 1) * dil / 2;
  a.pre_act = 1; a.pre_slope = 0.1f; a.lengths = len.data(); a.in_mask = 1;
  a.ep.out = out.data(); a.ep.out_bs = (long long)Cout * T; a.ep.out_mask = 1;
  if (mode == 0) { a.ep.mode = EPI_PLAIN; a.ep.act = 1; }
  if (mode == 1) { a.ep.mode = EPI_RESID; a.ep.resid = resid.data(); }
  if (mode == 2) { a.ep.mode = EPI_GATE; a.ep.H = Cout / 2; a.ep.out_bs = (long long)(Cout / 2) * T; }
  if (la) { a.la_len = la_len.data(); a.la_rate = 1; a.la_margin = 3; }
  p.wtc = reinterpret_cast<const float*>(packed);
  const int R = 128 + (K - 1) * dil;
  int nbuf = 2;
  const int nb_max = getenv("EMU_NB") ? atoi(getenv("EMU_NB")) : kTc16rNB;
  while (nbuf < nb_max && tc16r_smem_bytes(K, dil, N, KC, n_chunks, nbuf + 1) <= emu::kSmemBytes) ++nbuf;
  p.N = N; p.n_tiles = n_tiles; p.KC = KC; p.n_chunks = n_chunks; p.MB = 1; p.G = 1; p.n_bbuf = nbuf;
  p.R_pad = (R + 7) & ~7; p.tmem_cols = 512; p.l2_prefetch = 1;
  if (tc16r_smem_bytes(K, dil, N, KC, n_chunks, nbuf) > emu::kSmemBytes) { printf("smem over budget\n"); return 64; }
  printf("resident activations %zu B, %d weight slots, %d N tiles\n", (size_t)n_chunks * 2 * (KC / 8) * p.R_pad * 16, nbuf, n_tiles);
  unsigned long long n_mma = 0;
  if (mode == 0) emu::launch(conv1d_tc16r_kernel<EPI_PLAIN>, p, grid, kTc16rThreads, &n_mma);
  if (mode == 1) emu::launch(conv1d_tc16r_kernel<EPI_RESID>, p, grid, kTc16rThreads, &n_mma);
  if (mode == 2) emu::launch(conv1d_tc16r_kernel<EPI_GATE>, p, grid, kTc16rThreads, &n_mma);
  // reference
  double max_err = 0, sq = 0;
  long long cnt = 0;
  int untouched_ok = 1;
  const int group_rows = 128;
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t) {
      const bool skipped = la && ((long long)(t / group_rows) * group_rows >= (la_len[b] + 3));
      std::vector<double> y(Cout);
      for (int co = 0; co < Cout; ++co) {
        double s = bias[co];
        for (int ci = 0; ci < Cin; ++ci)
          for (int tap = 0; tap < K; ++tap) {
            const int ti = t + tap * dil - a.pad_left;
            if (ti < 0 || ti >= T || ti >= len[b]) continue;
            double xv = x[((size_t)b * Cin + ci) * T + ti];
            xv = xv > 0 ? xv : xv * 0.1;
            s += (double)w[((size_t)co * Cin + ci) * K + tap] * xv;
          }
        y[co] = s;
      }
      const double msk = t < len[b] ? 1.0 : 0.0;
      const int nout = mode == 2 ? Cout / 2 : Cout;
      for (int co = 0; co < nout; ++co) {
        double ref;
        if (mode == 0) ref = (y[co] > 0 ? y[co] : 0) * msk;
        else if (mode == 1) ref = y[co] + resid[((size_t)b * Cout + co) * T + t];
        else ref = tanh(y[2 * co]) / (1.0 + exp(-y[2 * co + 1]));
        const double got = out[((size_t)b * nout + co) * T + t];
        if (skipped) { untouched_ok = untouched_ok && (got == -555.0); continue; }
        const double e = fabs(ref - got);
        max_err = e > max_err ? e : max_err;
        sq += ref * ref;
        ++cnt;
      }
    }
  const double rms = sqrt(sq / (double)(cnt ? cnt : 1));
  printf("tc16r Cin=%d Cout=%d K=%d dil=%d B=%d T=%d N=%d KC=%d chunks=%d grid=%d mode=%d la=%d: mma=%llu max_err=%.3e rms=%.3e rel=%.3e untouched=%d\n",
         Cin, Cout, K, dil, B, T, N, KC, n_chunks, grid, mode, la, n_mma, max_err, rms, max_err / rms, untouched_ok);
  free(packed);
  return (max_err / rms < 2e-5 && untouched_ok) ? 0 : 2;
}
