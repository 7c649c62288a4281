// CPU check of fused_resblock2_kernel (wetts_b200/csrc/fused_rb_kernel.cuh) in the CTA emulator:
// the kernel source is compiled for the host and compared with a direct fp64 evaluation of
// ResBlock2 x nrb + MRF mean (decoders.py:205-214, :72-76).
//   usage: fused_rb_emu C B T grid [nrb] [ring slots: 4 | 6]
#define WETTS_EMULATE 1
#include <math.h>

#include <random>

#include "fused_rb_kernel.cuh"

using namespace wetts;

static double lrelu_d(double x, double s) { return x > 0 ? x : x * s; }

// y[co][t] = bias[co] + sum_{ci,tap} w[co][ci][tap] * lrelu(x)[ci][t + (tap - (k-1)/2)*d]   (zero padded)
static void conv_ref(const std::vector<double>& x, std::vector<double>& y, const std::vector<float>& w,
                     const std::vector<float>& bias, int C, int T, int k, int d, double slope) {
  for (int co = 0; co < C; ++co)
    for (int t = 0; t < T; ++t) {
      double s = bias[co];
      for (int ci = 0; ci < C; ++ci)
        for (int tap = 0; tap < k; ++tap) {
          const int ti = t + (tap - (k - 1) / 2) * d;
          if (ti >= 0 && ti < T) s += (double)w[((size_t)co * C + ci) * k + tap] * lrelu_d(x[(size_t)ci * T + ti], slope);
        }
      y[(size_t)co * T + t] = s;
    }
}

template <int C, int THREADS, int NB>
static int run(int B, int T, int grid, int nrb) {
  const int ks[3] = {3, 5, 7}, d1s[3] = {1, 2, 3}, d2s[3] = {2, 6, 12};
  std::mt19937 rng(1234 + C + T);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> x((size_t)B * C * T), out((size_t)B * C * T, -777.f);
  for (auto& v : x) v = nd(rng);
  FusedRbArgs a;
  a.B = B; a.T = T; a.nrb = nrb; a.slope = 0.1f; a.div = (float)nrb;
  std::vector<std::vector<float>> w1(nrb), w2(nrb), b1(nrb), b2(nrb);
  int Hmax = 0;
  size_t packed_floats = 0;
  for (int j = 0; j < nrb; ++j) {
    a.k[j] = ks[j]; a.d1[j] = d1s[j]; a.d2[j] = d2s[j];
    const int H = (d1s[j] + d2s[j]) * (ks[j] - 1) / 2;
    Hmax = H > Hmax ? H : Hmax;
    const float sc = 1.0f / sqrtf((float)(C * ks[j]));
    w1[j].resize((size_t)C * C * ks[j]); w2[j].resize((size_t)C * C * ks[j]);
    b1[j].resize(C); b2[j].resize(C);
    for (auto& v : w1[j]) v = nd(rng) * sc;
    for (auto& v : w2[j]) v = nd(rng) * sc;
    for (auto& v : b1[j]) v = nd(rng) * 0.1f;
    for (auto& v : b2[j]) v = nd(rng) * 0.1f;
    a.bias1[j] = b1[j].data(); a.bias2[j] = b2[j].data();
    packed_floats += 2 * (size_t)ks[j] * C * C * 2;
  }
  fused_rb_finalize_args(a, C);
  if (a.Rp < 128 + 2 * ((Hmax + 3) & ~3) || (a.Rp & 1) == 0) { printf("Rp mismatch\n"); return 1; }
  if ((size_t)a.nq * fused_rb_chunk_floats(C) != packed_floats) { printf("chunk accounting mismatch\n"); return 1; }
  if (NB == 6 && a.nq % 6 != 0) { printf("6-slot ring needs nq %% 6 == 0 (nq = %d)\n", a.nq); return 64; }
  if (fused_rb_smem_bytes(C, NB) > emu::kSmemBytes) { printf("smem over budget\n"); return 1; }
  // pack: conv order j0.c1, j0.c2, j1.c1, ... exactly as the engine does
  float* packed = (float*)aligned_alloc(128, packed_floats * 4);
  {
    size_t off = 0;
    for (int j = 0; j < nrb; ++j)
      for (int which = 0; which < 2; ++which) {
        const std::vector<float>& w = which ? w2[j] : w1[j];
        const long long n = (long long)ks[j] * C * C * 2;
        for (long long i = 0; i < n; ++i) {
          const FusedRbPackIdx ix = fused_rb_pack_index(i, C);
          const float v = w[((size_t)ix.co * C + ix.ci) * ks[j] + ix.tap];
          const float hi = tc::tf32_rna(v);
          packed[off + i] = ix.hl ? tc::tf32_rna(v - hi) : hi;
        }
        off += n;
      }
  }
  a.in = x.data(); a.out = out.data(); a.w = packed;
  a.smem_off = emu::kSmemBase;
  unsigned long long n_mma = 0;
  emu::launch(fused_resblock2_kernel<C, THREADS, 1, NB>, a, grid, THREADS, &n_mma);

  // reference
  double max_err = 0, sq = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<double> xb((size_t)C * T), acc((size_t)C * T, 0.0), y((size_t)C * T), x1((size_t)C * T);
    for (size_t i = 0; i < xb.size(); ++i) xb[i] = x[(size_t)b * C * T + i];
    for (int j = 0; j < nrb; ++j) {
      conv_ref(xb, y, w1[j], b1[j], C, T, ks[j], d1s[j], 0.1);
      for (size_t i = 0; i < y.size(); ++i) x1[i] = y[i] + xb[i];
      conv_ref(x1, y, w2[j], b2[j], C, T, ks[j], d2s[j], 0.1);
      for (size_t i = 0; i < y.size(); ++i) acc[i] += y[i] + x1[i];
    }
    for (size_t i = 0; i < acc.size(); ++i) {
      const double ref = acc[i] / nrb, got = out[(size_t)b * C * T + i];
      const double e = fabs(ref - got);
      max_err = e > max_err ? e : max_err;
      sq += ref * ref;
    }
  }
  const double rms = sqrt(sq / ((double)B * C * T));
  printf("C=%d ring=%d B=%d T=%d grid=%d nrb=%d: mma=%llu max_err=%.3e rms=%.3e rel=%.3e\n", C, NB, B, T, grid, nrb, n_mma, max_err, rms,
         max_err / rms);
  free(packed);
  return (max_err / rms < 2e-5) ? 0 : 2;
}

int main(int argc, char** argv) {
  if (argc < 5) { printf("usage: %s C B T grid [nrb] [ring]\n", argv[0]); return 64; }
  const int C = atoi(argv[1]), B = atoi(argv[2]), T = atoi(argv[3]), grid = atoi(argv[4]);
  const int nrb = argc > 5 ? atoi(argv[5]) : 3;
  const int ring = argc > 6 ? atoi(argv[6]) : 4;
  if (C == 32 && ring == 4) return run<32, 256, 4>(B, T, grid, nrb);
  if (C == 32 && ring == 6) return run<32, 256, 6>(B, T, grid, nrb);
  if (C == 64 && ring == 4) return run<64, 512, 4>(B, T, grid, nrb);
  if (C == 64 && ring == 6) return run<64, 512, 6>(B, T, grid, nrb);
  return 64;
}
