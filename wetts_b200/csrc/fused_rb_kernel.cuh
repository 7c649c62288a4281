// Fused HiFi-GAN MRF stage for ResBlock2 generators (decoders.py:205-214, :72-76), sm_100a tcgen05.
//
//   out = (1/nrb) * sum_j  rb_j(x),     rb_j(x):  x1 = conv_{k_j,d1_j}(lrelu(x)) + x
//                                                 x2 = conv_{k_j,d2_j}(lrelu(x1)) + x1
//
// One launch replaces the 2*nrb conv launches of a stage: x is read from HBM (L2 for the repeats),
// x1 never leaves the SM, the MRF sum lives in registers and `out` is written once.
//
// Work item = (utterance b, 128 output samples).  For resblock j with halos h1 = d1(k-1)/2,
// h2 = d2(k-1)/2, H = h1+h2 the CTA stages R = 128 + 2H input rows as the tcgen05 A operand
// (M = time rows, K = channels; no-swizzle K-major canonical layout, hi/lo TF32 split, lrelu fused):
//   element (row r, channel c) at A + (c/4)*Rp*16 + r*16 + (c%4)*4     (lo tile at + a_half)
// so a conv tap is a row shift of the descriptor start address.  Row r <-> sample t0 - H + r.
//   conv1: two M = 128 blocks starting at rows 0 and 2*h2 produce x1 on rows [h1, R-h1); the epilogue
//          writes lrelu(x1) (hi/lo) IN PLACE over lrelu(x) -- legal because every conv1 MMA has
//          completed (acc1 barrier) before the first row is overwritten; rows outside [0,T) are zeroed
//          (the reference's zero padding of conv2's input).
//   conv2: one block starting at row h1 produces the 128 outputs; residual x1 comes from the same tile.
// The residual is recovered from the staged lrelu value (hi + lo, inverse lrelu): relative deviation
// <= 2^-22, the same order as the 3xTF32 product error.
//
// fp32 accuracy: 3xTF32 in the two-MMA form of tc_mma_tf32_split2 (weights stored [hi | lo] along N; the
// accumulator of a block is [hi*hi | hi*lo + lo*hi], summed in the epilogue).
//
// Weights stream through a 4-slot shared-memory ring of 32-input-channel chunks (one tap, hi+lo,
// C/32 chunks per tap) fetched by cp.async.bulk three chunks ahead of the MMAs; the chunk sequence is
// identical for every item so the ring never drains between items.  The next tile's activations are
// prefetched into registers while the tensor pipe works.
//
// This file contains no PTX: everything hardware specific is in tc_prims.cuh, and the same source runs
// in the host CTA emulator (tests/emu).
#pragma once
#include "fused_rb_args.h"
#include "tc_prims.cuh"

namespace wetts {

template <int C, int THREADS, int MINB>
WETTS_GLOBAL void WETTS_LAUNCH_BOUNDS(THREADS, MINB) fused_resblock2_kernel(const FusedRbArgs p) {
  using namespace tc;
  static_assert(C == 32 || C == 64, "channel count");
  static_assert(THREADS == 8 * C, "8 warps for C = 32, 16 warps for C = 64");
  constexpr int N = C;
  constexpr int NB = kFusedRbRing, PD = kFusedRbAhead, NU = kFusedRbUnits;
  constexpr int KH = C / 32;
  constexpr uint32_t CHUNK_BYTES = 8u * 2u * N * 16u;     // [8 k-groups][hi|lo][N][4 floats]
  constexpr uint32_t TMEM_COLS = (6 * N <= 256) ? 256u : 512u;   // 3 accumulator blocks x [hi*hi | small terms]
  constexpr int CG = C / 4;

  WETTS_SMEM_DECL(smem);
  const int tid = WETTS_TID, lane = tid & 31;
  const int warp = (int)warp_uniform((uint32_t)(tid >> 5));
  const int T = p.T, Rp = p.Rp, nrb = p.nrb;
  const uint32_t a_half = (uint32_t)C * (uint32_t)Rp * 4u;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 120);
  // weight ring FIRST: its shared-memory address is then a compile-time offset (not a function of the launch
  // parameter Rp), which lets the compiler rematerialise the B descriptors in uniform registers
  uint8_t* Ahi = smem + 128 + NB * CHUNK_BYTES;
  const uint32_t bar_full = smem_u32(&bars[0]);        // [NB] TMA -> MMA: weight chunk landed
  const uint32_t bar_empty = smem_u32(&bars[NB]);      // [NB] MMA -> TMA: weight slot reusable
  const uint32_t bar_acc1 = smem_u32(&bars[2 * NB]);   //      conv1 accumulators complete
  const uint32_t bar_acc2 = smem_u32(&bars[2 * NB + 1]);
  const uint32_t A_addr = smem_u32(Ahi);
  const uint32_t ring_addr = smem_u32(smem + 128);

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2 * NB + 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init_fence();
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = warp_uniform(*tmem_slot);

  const int n_ttiles = (T + 127) / 128;
  const int n_items = p.B * n_ttiles;
  const int my_items = (WETTS_BID < n_items) ? (n_items - WETTS_BID + WETTS_NBLK - 1) / WETTS_NBLK : 0;
  const long long bs = (long long)C * T;
  const float inv_slope = 1.0f / p.slope;

  // ---------------------------------------------------------------- activation prefetch / staging
  float pf[NU][4];
  auto prefetch = [&](int item, int j) {
    const int b = item / n_ttiles;
    const int t0 = (item - b * n_ttiles) * 128;
    const int H = (p.d1[j] + p.d2[j]) * (p.k[j] - 1) / 2;
    const int R = 128 + 2 * H;
    const float* in_b = p.in + (long long)b * bs;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int u = tid + i * THREADS;
      const int cg = u / R, r = u - cg * R;
      const int t = t0 - H + r;
      const bool ok = (cg < CG) && (t >= 0) && (t < T);
      const float* src = in_b + (long long)(4 * cg) * T + t;
#pragma unroll
      for (int e = 0; e < 4; ++e) pf[i][e] = ok ? ldg(src + (long long)e * T) : 0.f;
    }
  };
  auto split_store = [&](int cg, int r, float y0, float y1, float y2, float y3) {
    float4 hi, lo;
    hi.x = tf32_rna(y0); lo.x = tf32_rna(y0 - hi.x);
    hi.y = tf32_rna(y1); lo.y = tf32_rna(y1 - hi.y);
    hi.z = tf32_rna(y2); lo.z = tf32_rna(y2 - hi.z);
    hi.w = tf32_rna(y3); lo.w = tf32_rna(y3 - hi.w);
    uint8_t* dst = Ahi + ((size_t)cg * Rp + r) * 16;
    *reinterpret_cast<float4*>(dst) = hi;
    *reinterpret_cast<float4*>(dst + a_half) = lo;
  };
  auto lrelu = [&](float x) { return x > 0.f ? x : x * p.slope; };
  auto stage = [&](int j) {
    const int H = (p.d1[j] + p.d2[j]) * (p.k[j] - 1) / 2;
    const int R = 128 + 2 * H;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int u = tid + i * THREADS;
      const int cg = u / R, r = u - cg * R;
      if (cg < CG) split_store(cg, r, lrelu(pf[i][0]), lrelu(pf[i][1]), lrelu(pf[i][2]), lrelu(pf[i][3]));
    }
  };
  // pre-activation value recovered from the staged hi/lo pair of (row r, channel group cg)
  auto staged_value = [&](int cg, int r, float* x) {
    const uint8_t* src = Ahi + ((size_t)cg * Rp + r) * 16;
    const float4 hi = *reinterpret_cast<const float4*>(src);
    const float4 lo = *reinterpret_cast<const float4*>(src + a_half);
    const float y0 = hi.x + lo.x, y1 = hi.y + lo.y, y2 = hi.z + lo.z, y3 = hi.w + lo.w;
    x[0] = y0 > 0.f ? y0 : y0 * inv_slope;
    x[1] = y1 > 0.f ? y1 : y1 * inv_slope;
    x[2] = y2 > 0.f ? y2 : y2 * inv_slope;
    x[3] = y3 > 0.f ? y3 : y3 * inv_slope;
  };

  // ---------------------------------------------------------------- weight ring (warp 0, warp-uniform)
  // Chunk number g = it*nq + q (q = index within the item) is pure arithmetic on loop counters and kernel
  // parameters, so slot = g % NB, phase = g / NB and every descriptor derived from them stay in uniform
  // registers: the tcgen05.mma operands need no R2UR (the issue loop is otherwise ~90 cycles per MMA).
  static_assert((NB & (NB - 1)) == 0, "ring size must be a power of two");
  constexpr uint32_t LOG_NB = (NB == 2) ? 1u : (NB == 4) ? 2u : 3u;
  const uint32_t nq = (uint32_t)p.nq;
  auto produce = [&](uint32_t it_p, uint32_t q_p) {      // request chunk q_p of this CTA's it_p-th item
    if (q_p >= nq) { q_p -= nq; it_p += 1; }
    if (it_p >= (uint32_t)my_items) return;
    const uint32_t g = it_p * nq + q_p;
    const uint32_t slot = g & (uint32_t)(NB - 1), use = g >> LOG_NB;
    if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);
    if (elect_one()) {
      mbar_expect_tx(bar_full + 8 * slot, CHUNK_BYTES);
      bulk_g2s(ring_addr + slot * CHUNK_BYTES, reinterpret_cast<const uint8_t*>(p.w) + (size_t)q_p * CHUNK_BYTES,
               CHUNK_BYTES, bar_full + 8 * slot);
    }
    warp_sync();
  };
  if (warp == 0)
    for (int i = 0; i < PD; ++i) produce(0u, (uint32_t)i);

  const uint32_t idesc_n = idesc_tf32_m128(N), idesc_2n = idesc_tf32_m128(2 * N);
  const uint64_t adesc0 = make_desc(A_addr, (uint32_t)Rp * 16u, 128u);
  const uint64_t bdesc0 = make_desc(ring_addr, (uint32_t)(2 * N) * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  const uint32_t a_lo_delta = a_half >> 4;

  // One conv on the tensor pipe: for every tap and 32-channel slice, multiply the weight chunk with
  // `nblk` 128-row blocks of the activation tile (block m starts at row row0 + m*row_step + tap*dil).
  auto run_conv = [&](uint32_t it, uint32_t qbase, int k, int dil, int nblk, int row0, int row_step, uint32_t d_col0,
                      uint32_t done_bar) {
    for (int tap = 0; tap < k; ++tap) {
      for (int kh = 0; kh < KH; ++kh) {
        const uint32_t q = qbase + (uint32_t)(tap * KH + kh);
        const uint32_t g = it * nq + q;
        const uint32_t slot = g & (uint32_t)(NB - 1), par = (g >> LOG_NB) & 1u;
        mbar_wait(bar_full + 8 * slot, par);
        tc_fence_after();
        const uint32_t b0 = blo0 + slot * (CHUNK_BYTES >> 4);
        const uint32_t first = (tap == 0 && kh == 0) ? 0u : 1u;
        for (int m = 0; m < nblk; ++m) {
          const uint32_t a0 = alo0 + (uint32_t)((kh * 8) * Rp + row0 + m * row_step + tap * dil);
          const uint32_t d_tmem = tmem_base + d_col0 + (uint32_t)(m * 2 * N);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t al = a0 + (uint32_t)(kk * 2 * Rp), bl = b0 + (uint32_t)(kk * 2 * 2 * N);
            tc_mma_tf32_split2(d_tmem, d_tmem + (uint32_t)N, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + a_lo_delta),
                               desc_with_lo(bdesc0, bl), idesc_2n, idesc_n, (kk == 0) ? first : 1u);
          }
        }
        if (elect_one()) tc_commit(bar_empty + 8 * slot);
        warp_sync();
        produce(it, q + (uint32_t)PD);
      }
    }
    if (elect_one()) tc_commit(done_bar);
    warp_sync();
  };

  // ---------------------------------------------------------------- main loop
  const int q4 = warp & 3, grp = warp >> 2;
  const int row_i = 32 * q4 + lane;                   // TMEM lane = row of the 128-row block
  const uint32_t lane_sel = (uint32_t)(32 * q4) << 16;
  uint32_t rb_count = 0;
  if (my_items > 0) prefetch(WETTS_BID, 0);

  for (int it = 0; it < my_items; ++it) {
    const int item = WETTS_BID + it * WETTS_NBLK;
    const int b = item / n_ttiles;
    const int t0 = (item - b * n_ttiles) * 128;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;

    for (int j = 0; j < nrb; ++j) {
      const int k = p.k[j], d1 = p.d1[j], d2 = p.d2[j];
      const int h1 = d1 * (k - 1) / 2, h2 = d2 * (k - 1) / 2, H = h1 + h2;

      stage(j);
      fence_async_smem();
      tc_fence_before();
      cta_sync();
      tc_fence_after();
      // the next tile's input travels while the tensor pipe works on this one
      if (j + 1 < nrb) prefetch(item, j + 1);
      else if (it + 1 < my_items) prefetch(item + WETTS_NBLK, 0);

      // ---- conv1: x1 on rows [h1, R - h1)
      if (warp == 0) run_conv((uint32_t)it, (uint32_t)p.qoff[2 * j], k, d1, 2, 0, 2 * h2, 0u, bar_acc1);
      mbar_wait(bar_acc1, rb_count & 1);
      tc_fence_after();
      {
        const int mb = grp & 1, cbase = 32 * (grp >> 1);
        const int r1 = (mb ? 2 * h2 : 0) + row_i + h1;
        const bool active = (mb == 0) || (row_i >= 128 - 2 * h2);   // block 1 only adds the rows block 0 lacks
        const int t = t0 - H + r1;
        const bool inside = (t >= 0) && (t < T);
        const float* bias = p.bias1[j];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = cbase + 16 * half;
          float v[16], vs[16];
          tmem_ld16(tmem_base + lane_sel + (uint32_t)(mb * 2 * N + c0), v);
          tmem_ld16(tmem_base + lane_sel + (uint32_t)(mb * 2 * N + N + c0), vs);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += vs[i];
          if (active) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const int cg = (c0 >> 2) + g4;
              float x[4], y[4];
              staged_value(cg, r1, x);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x1 = inside ? (v[4 * g4 + e] + ldg(bias + c0 + 4 * g4 + e)) + x[e] : 0.f;
                y[e] = lrelu(x1);
              }
              split_store(cg, r1, y[0], y[1], y[2], y[3]);
            }
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      cta_sync();
      tc_fence_after();

      // ---- conv2: the 128 outputs of this item
      if (warp == 0) run_conv((uint32_t)it, (uint32_t)p.qoff[2 * j + 1], k, d2, 1, h1, 0, (uint32_t)(4 * N), bar_acc2);
      mbar_wait(bar_acc2, rb_count & 1);
      tc_fence_after();
      {
        const int c0 = 16 * grp;
        const float* bias = p.bias2[j];
        float v[16], vs[16];
        tmem_ld16(tmem_base + lane_sel + (uint32_t)(4 * N + c0), v);
        tmem_ld16(tmem_base + lane_sel + (uint32_t)(5 * N + c0), vs);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += vs[i];
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          float x[4];
          staged_value((c0 >> 2) + g4, H + row_i, x);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[4 * g4 + e] += (v[4 * g4 + e] + ldg(bias + c0 + 4 * g4 + e)) + x[e];
        }
        if (j == nrb - 1) {
          const int t = t0 + row_i;
          if (t < T) {
            float* op = p.out + (long long)b * bs + (long long)c0 * T + t;
#pragma unroll
            for (int i = 0; i < 16; ++i) op[(long long)i * T] = (nrb > 1) ? acc[i] / p.div : acc[i];
          }
        }
      }
      rb_count += 1;
      // every thread is done with the activation tile and with TMEM before either is overwritten
      tc_fence_before();
      cta_sync();
      tc_fence_after();
    }
  }
  cta_sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace wetts
