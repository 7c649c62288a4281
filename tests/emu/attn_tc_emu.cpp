// CPU check of rel_attention_tc_kernel (wetts_b200/csrc/attn_tc_kernel.cuh) in the CTA emulator against an fp64
// evaluation of the reference's windowed relative-position attention (attentions.py:232-282).
//   usage: attn_tc_emu B T [len0 len1 ...]
#define WETTS_EMULATE 1
#include <stdlib.h>
#include <math.h>

#include <random>

#include "attn_tc_kernel.cuh"

using namespace wetts;

int main(int argc, char** argv) {
  if (argc < 3) { printf("usage: %s B T [lengths...]\n", argv[0]); return 64; }
  const int B = atoi(argv[1]), T = atoi(argv[2]);
  const int H = 2, DK = kAttnTcDk, C = H * DK, W = 4, NREL = 9;
  if (T > kAttnTcMaxT) { printf("T too large\n"); return 64; }
  std::vector<long long> len(B, T);
  for (int b = 0; b < B && 3 + b < argc; ++b) len[b] = atoi(argv[3 + b]);
  std::mt19937 rng(99 + T);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> qkv((size_t)B * 3 * C * T), ek((size_t)NREL * DK), ev((size_t)NREL * DK), out((size_t)B * C * T, -7.f);
  for (auto& v : qkv) v = nd(rng);
  for (auto& v : ek) v = nd(rng) * 0.1f;
  for (auto& v : ev) v = nd(rng) * 0.1f;
  AttnTcArgs a;
  a.qkv = qkv.data(); a.emb_k = ek.data(); a.emb_v = ev.data(); a.lengths = len.data(); a.out = out.data();
  a.B = B; a.C = C; a.T = T; a.n_heads = H; a.window = W; a.smem_off = emu::kSmemBase;
  if (kAttnTcSmem > emu::kSmemBytes) { printf("smem over budget\n"); return 1; }
  unsigned long long n_mma = 0;
  // EMU_ATTN_SHARED=1: the two-CTAs-per-SM variant (shared memory used twice, O in the TMEM columns of S)
  const bool shared = getenv("EMU_ATTN_SHARED") && atoi(getenv("EMU_ATTN_SHARED"));
  if (shared) emu::launch(rel_attention_tc_kernel<true>, a, B * H, kAttnTcThreads, &n_mma);
  else emu::launch(rel_attention_tc_kernel<false>, a, B * H, kAttnTcThreads, &n_mma);
  double max_err = 0, sq = 0;
  const double scale = 1.0 / sqrt((double)DK);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      const float* q = qkv.data() + ((size_t)b * 3 * C + h * DK) * T;
      const float* k = q + (size_t)C * T;
      const float* v = k + (size_t)C * T;
      for (int i = 0; i < T; ++i) {
        std::vector<double> s(T), pr(T);
        double mx = -1e300;
        for (int j = 0; j < T; ++j) {
          double x = 0;
          for (int d = 0; d < DK; ++d) x += (double)q[(size_t)d * T + i] * scale * (double)k[(size_t)d * T + j];
          const int r = j - i + W;
          if (r >= 0 && r < NREL)
            for (int d = 0; d < DK; ++d) x += (double)q[(size_t)d * T + i] * scale * (double)ek[(size_t)r * DK + d];
          if (i >= len[b] || j >= len[b]) x = -1e4;
          s[j] = x;
          mx = x > mx ? x : mx;
        }
        double sum = 0;
        for (int j = 0; j < T; ++j) { pr[j] = exp(s[j] - mx); sum += pr[j]; }
        for (int d = 0; d < DK; ++d) {
          double o = 0;
          for (int j = 0; j < T; ++j) o += pr[j] / sum * (double)v[(size_t)d * T + j];
          for (int r = 0; r < NREL; ++r) {
            const int j = i + r - W;
            if (j >= 0 && j < T) o += pr[j] / sum * (double)ev[(size_t)r * DK + d];
          }
          const double got = out[((size_t)b * C + h * DK + d) * T + i];
          const double e = fabs(o - got);
          max_err = e > max_err ? e : max_err;
          sq += o * o;
        }
      }
    }
  const double rms = sqrt(sq / ((double)B * C * T));
  printf("attn_tc B=%d T=%d: mma=%llu max_err=%.3e rms=%.3e rel=%.3e\n", B, T, n_mma, max_err, rms, max_err / rms);
  return max_err / rms < 2e-5 ? 0 : 2;
}
