// Arguments and packing rules of the f16-split fused HiFi-GAN MRF stage kernel (fused_mrf16_kernel.cuh).
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

struct int2_t { int x, y; };
constexpr int kMrfMaxRb = 3;       // resblocks per stage (MRF branches)
constexpr int kMrfMaxConv = 6;     // convs per resblock: ResBlock2 = 2, ResBlock1 = 6 (c1_d0, c2, c1_d1, c2, c1_d2, c2)
constexpr int kMrfProfPhases = 12;

struct FusedMrfArgs {
  const float* in = nullptr;    // [B][C][T]
  float* out = nullptr;         // [B][C][T]
  const void* w = nullptr;      // packed f16 chunk sequence of one item (fused_mrf16_pack_index)
  const float* bias[kMrfMaxRb][kMrfMaxConv] = {};
  int B = 0, T = 0, nrb = 0;
  int type = 2;                 // 1: ResBlock1 (decoders.py:157-170), 2: ResBlock2 (decoders.py:205-214)
  int nconv = 2;                // convs per resblock
  int k[kMrfMaxRb] = {0, 0, 0};
  int dil[kMrfMaxRb][kMrfMaxConv] = {};
  int qoff[kMrfMaxRb][kMrfMaxConv] = {};   // first weight chunk (within the item) of conv c of resblock j
  int nq = 0;                   // weight chunks per item = sum_j nconv * k_j * (C/32)
  int nq_ring = 0;              // nq / ring slots when the ring size divides nq, else 0
  uint32_t smem_off = 0;        // CTA-local shared-window offset of the dynamic shared memory base
  float slope = 0.1f;
  float div = 1.f;
  // length-aware mode (optional): item_map[i] = {utterance, first sample} of work item i and *n_items_dev = number
  // of items (built on the device from the utterance lengths, launch_mrf_item_map); nullptr = every 128-sample
  // tile of every utterance, in (b, tile) order
  const int2_t* item_map = nullptr;
  const int* n_items_dev = nullptr;
  // start-up stagger (cycles) of the k-th co-resident CTA of an SM (k = block index / n_sm): co-resident CTAs that
  // start together fall into a convoy (all in their MMA phase, then all in their SIMT phase); 0 = off
  int stagger = 0, n_sm = 1;
  long long* prof = nullptr;    // profiling instantiation only
};

// floats-equivalent: one weight chunk = one tap x 32 input channels: [4 k-groups of 8 ch][hi | lo' rows: 2N][8 halfs]
constexpr int fused_mrf16_chunk_bytes(int C) { return 4 * 2 * C * 16; }
inline int fused_mrf16_ring_slots(int nq) { return (nq % 6 == 0) ? 6 : 4; }
// activation tile: [C/8 groups][Rp rows][8 halfs], hi tile then lo' tile
inline size_t fused_mrf16_tile_bytes(int C, int Rp) { return 2 * (size_t)(C / 8) * Rp * 16; }
inline size_t fused_mrf16_smem_bytes(int C, int ring, int Rp, int tiles) {
  return 128 + (size_t)ring * fused_mrf16_chunk_bytes(C) + (size_t)kMrfMaxRb * kMrfMaxConv * C * 4 +
         (size_t)tiles * fused_mrf16_tile_bytes(C, Rp);
}
inline size_t fused_mrf16_conv_halfs(int C, int K) { return (size_t)K * C * C * 2; }   // hi + lo'

// total halo of resblock j (rows on each side of the 128 output rows)
WETTS_HD inline int fused_mrf16_halo(const FusedMrfArgs& a, int j) {
  int H = 0;
  for (int c = 0; c < a.nconv; ++c) H += a.dil[j][c] * (a.k[j] - 1) / 2;
  return H;
}

// Derived launch fields (nq, qoff, nq_ring) from (nrb, nconv, k).
inline void fused_mrf16_finalize_args(FusedMrfArgs& a, int C) {
  a.nq = 0;
  for (int j = 0; j < a.nrb; ++j)
    for (int c = 0; c < a.nconv; ++c) {
      a.qoff[j][c] = a.nq;
      a.nq += a.k[j] * (C / 32);
    }
  a.nq_ring = (a.nq % 6 == 0) ? a.nq / 6 : 0;
}

// Packed weight element i (in halfs) of a conv [C][C][k] (chunk order: tap, 32-channel slice): source coordinates.
// i indexes [tap][kh][kg 0..3][hl][n][e 0..7];  ci = kh*32 + kg*8 + e.
struct FusedMrfPackIdx { int tap, hl, co, ci; };
WETTS_HD inline FusedMrfPackIdx fused_mrf16_pack_index(long long i, int C) {
  FusedMrfPackIdx r;
  const int e = (int)(i % 8); i /= 8;
  const int n = (int)(i % C); i /= C;
  r.hl = (int)(i % 2); i /= 2;
  const int kg = (int)(i % 4); i /= 4;
  const int kh = (int)(i % (C / 32)); i /= (C / 32);
  r.tap = (int)i;
  r.co = n;
  r.ci = kh * 32 + kg * 8 + e;
  return r;
}

}  // namespace wetts
