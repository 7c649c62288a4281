// Arguments and packing rules of the fused ResBlock2/MRF stage kernel (fused_rb_kernel.cuh).
// Plain C++ (no CUDA dependency): shared by the kernel, the engine and the host CTA emulator.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WETTS_HD __host__ __device__
#else
#define WETTS_HD
#endif

namespace wetts {

struct FusedRbArgs {
  const float* in = nullptr;   // [B][C][T]
  float* out = nullptr;        // [B][C][T]
  const float* w = nullptr;    // packed chunk sequence of one item (fused_rb_pack_index)
  const float* bias1[3] = {nullptr, nullptr, nullptr};
  const float* bias2[3] = {nullptr, nullptr, nullptr};
  int B = 0, T = 0, nrb = 0;
  int k[3] = {0, 0, 0}, d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0};
  int Rp = 0;                  // row pitch of a channel group in the activation buffer (= kFusedRbPitch; informational)
  int nq = 0;                  // weight chunks per item = sum_j 2*k_j*(C/32)
  int nq_ring = 0;             // nq / ring slots when the ring size divides nq (6-slot ring), else 0
  int qoff[6] = {0, 0, 0, 0, 0, 0};  // first chunk (within the item) of conv 2*j + {0: conv1, 1: conv2}
  uint32_t smem_off = 0;       // CTA-local shared-window offset of the dynamic shared memory base (see the kernel)
  float slope = 0.1f;
  float div = 1.f;
  long long* prof = nullptr;   // profiling instantiation only: [grid][2][kFusedRbProfPhases] cycle sums
};
constexpr int kFusedRbProfPhases = 15;

constexpr int kFusedRbPitch = 225;   // rows per 4-channel group of the activation tile (odd; >= 128 + 2*48 halo rows)
constexpr int kFusedRbUnits = 2;     // (4 rows x 4 channels) staging units per thread  ->  R <= 256 rows

// floats of one weight chunk: [8 k-groups][hi|lo][N][4]  (hi and lo rows adjacent: one 2N-row operand)
constexpr int fused_rb_chunk_floats(int C) { return 2 * 8 * C * 4; }
// weight ring slots: 6 when the chunk count per item is a multiple of 6 (slot index then depends on the chunk's
// position in the item only), else 4 (slot and phase are bit fields of the global chunk number)
inline int fused_rb_ring_slots(int nq) { return (nq % 6 == 0) ? 6 : 4; }
inline size_t fused_rb_smem_bytes(int C, int ring, int Rp = kFusedRbPitch) {
  return 128 + (size_t)ring * fused_rb_chunk_floats(C) * 4 + 6 * (size_t)C * 4 + 2 * (size_t)C * Rp * 4;
}
// Derived launch fields (nq, qoff, nq_ring, Rp) from (nrb, k).
inline void fused_rb_finalize_args(FusedRbArgs& a, int C) {
  a.nq = 0;
  for (int j = 0; j < a.nrb; ++j) {
    a.qoff[2 * j] = a.nq;
    a.nq += a.k[j] * (C / 32);
    a.qoff[2 * j + 1] = a.nq;
    a.nq += a.k[j] * (C / 32);
  }
  a.nq_ring = (a.nq % 6 == 0) ? a.nq / 6 : 0;
  a.Rp = kFusedRbPitch;   // compile-time pitch; fused_rb_supported() guarantees 128 + 2*Hp_max <= pitch
}

// Packed weight element i of a conv [C][C][k] (chunk order: tap, 32-channel slice): returns the source
// coordinates.  i indexes [tap][kh][kg][hl][n][e].
struct FusedRbPackIdx { int tap, hl, co, ci; };
WETTS_HD inline FusedRbPackIdx fused_rb_pack_index(long long i, int C) {
  FusedRbPackIdx r;
  const int e = (int)(i % 4); i /= 4;
  const int n = (int)(i % C); i /= C;
  r.hl = (int)(i % 2); i /= 2;
  const int kg = (int)(i % 8); i /= 8;
  const int kh = (int)(i % (C / 32)); i /= (C / 32);
  r.tap = (int)i;
  r.co = n;
  r.ci = kh * 32 + kg * 4 + e;
  return r;
}

}  // namespace wetts
