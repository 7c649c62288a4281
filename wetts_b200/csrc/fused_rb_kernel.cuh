// Fused HiFi-GAN MRF stage for ResBlock2 generators (decoders.py:205-214, :72-76), sm_100a tcgen05.
//
//   out = (1/nrb) * sum_j  rb_j(x),     rb_j(x):  x1 = conv_{k_j,d1_j}(lrelu(x)) + x
//                                                 x2 = conv_{k_j,d2_j}(lrelu(x1)) + x1
//
// One launch replaces the 2*nrb conv launches of a stage: x is read from HBM (L2 for the repeats),
// x1 never leaves the SM, the MRF sum accumulates on chip (conv terms in TMEM across the resblocks,
// residual + bias terms in registers) and `out` is written once.
//
// Work item = (utterance b, 128 output samples).  For resblock j with halos h1 = d1(k-1)/2,
// h2 = d2(k-1)/2, H = h1+h2, Hp = H rounded up to 4 (16 B aligned loads), dl = Hp - H, the CTA stages
// R = 128 + 2Hp input rows as the tcgen05 A operand (M = time rows, K = channels; no-swizzle K-major
// canonical layout, hi/lo TF32 split, lrelu fused):
//   element (row r, channel c) at A + (c/4)*Rp*16 + r*16 + (c%4)*4     (lo tile at + a_half)
// so a conv tap is a row shift of the descriptor start address.  Row r <-> sample t0 - Hp + r.  Rp is odd:
// the 4-channel groups then start 16 B apart modulo 128 B, so the staging stores (lanes across channel groups)
// are bank-conflict free, like the epilogue accesses (lanes across rows).
//   conv1: two M = 128 blocks starting at rows dl and dl + 2*h2 produce x1 on rows [dl+h1, R-dl-h1); the epilogue
//          writes lrelu(x1) (hi/lo) IN PLACE over lrelu(x) -- legal because every conv1 MMA has
//          completed (acc1 barrier) before the first row is overwritten; rows outside [0,T) are zeroed
//          (the reference's zero padding of conv2's input).
//   conv2: one block starting at row dl + h1 produces the 128 outputs; residual x1 comes from the same tile.
// The residual is recovered from the staged lrelu value (hi + lo, inverse lrelu): relative deviation
// <= 2^-22, the same order as the 3xTF32 product error.
//
// fp32 accuracy: 3xTF32 in the two-MMA form of tc_mma_tf32_split2 (weights stored [hi | lo] along N; the
// accumulator of a block is [hi*hi | hi*lo + lo*hi], summed in the epilogue).
//
// Weights stream through a 4- or 6-slot shared-memory ring of 32-input-channel chunks (one tap, hi+lo,
// C/32 chunks per tap): a slot is refilled by cp.async.bulk (L2 evict_last) the moment the MMAs that read it
// complete; the chunk sequence is identical for every item so the ring never drains between items.  Warp 0 only issues MMAs (the tensor
// pipe's instruction queue is shallow, so every cycle the issuer spends elsewhere is a pipe bubble); warp 1
// runs the weight producer during the MMA phases.  The next tile's activations are prefetched into registers
// (16 B loads, 4 rows x 4 channels per unit) while the tensor pipe works.
//
// This file contains no PTX: everything hardware specific is in tc_prims.cuh, and the same source runs
// in the host CTA emulator (tests/emu).
#pragma once
#include "fused_rb_args.h"
#include "tc_prims.cuh"

namespace wetts {

// PROFILE = true adds clock64 phase timers (thread 0 = the MMA issuer, thread 32 = the producer warp),
// summed per CTA into p.prof[cta][2][kFusedRbProfPhases]; used by tools/, never by the product path.
// NB = weight ring slots: 4 (any chunk count) or 6 (needs nq % 6 == 0; 50 % more look-ahead).
template <int C, int THREADS, int MINB, int NB = 4, bool PROFILE = false>
WETTS_GLOBAL void WETTS_LAUNCH_BOUNDS(THREADS, MINB) fused_resblock2_kernel(const FusedRbArgs p) {
  using namespace tc;
  static_assert(C == 32 || C == 64, "channel count");
  static_assert(THREADS == 8 * C, "8 warps for C = 32, 16 warps for C = 64");
  constexpr int N = C;
  static_assert(NB == 4 || NB == 6, "ring size");
  constexpr int PD = NB, NU = kFusedRbUnits;   // a slot is refilled the moment its chunk completes
  constexpr int KH = C / 32;
  constexpr uint32_t CHUNK_BYTES = 8u * 2u * N * 16u;     // [8 k-groups][hi|lo][N][4 floats]
  constexpr uint32_t TMEM_COLS = (6 * N <= 256) ? 256u : 512u;   // 3 accumulator blocks x [hi*hi | small terms]
  constexpr int CG = C / 4;
  constexpr int LOG_CG = (CG == 8) ? 3 : 4;
  // With 16 warps the issuer (warp 0) and the weight producer (warp 1) do not stage activations: their global
  // loads would sit between them and the tensor pipe.  The other 14 warps cover the tile exactly (448 x 2 units).
  constexpr int SW0 = (C == 64) ? 2 : 0;                 // first staging warp
  constexpr int STHREADS = THREADS - 32 * SW0;
  static_assert(NU * STHREADS >= CG * (kFusedRbPitch / 4), "staging units do not cover the activation tile");

  WETTS_SMEM_DECL(smem);
  const int tid = WETTS_TID, lane = tid & 31;
  const int warp = (int)uniform_bits((uint32_t)(tid >> 5), 0, 4);
  const int T = p.T, nrb = p.nrb;
  // compile-time row pitch: every descriptor offset (k-step, hi/lo tile) is then an immediate, which keeps
  // the number of live uniform registers in the MMA issue loop small
  constexpr int Rp = kFusedRbPitch;
  constexpr uint32_t a_half = (uint32_t)C * (uint32_t)Rp * 4u;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 120);
  // weight ring FIRST: its shared-memory address is then a compile-time offset (not a function of the launch
  // parameter Rp), which lets the compiler rematerialise the B descriptors in uniform registers
  float* bias_s = reinterpret_cast<float*>(smem + 128 + NB * CHUNK_BYTES);   // [6][C]: conv 2j (+1) of resblock j
  uint8_t* Ahi = smem + 128 + NB * CHUNK_BYTES + 6 * C * 4;
  const uint32_t bar_full = smem_u32(&bars[0]);        // [NB] TMA -> MMA: weight chunk landed
  const uint32_t bar_empty = smem_u32(&bars[NB]);      // [NB] MMA -> TMA: weight slot reusable
  const uint32_t bar_acc1 = smem_u32(&bars[2 * NB]);   //      conv1 accumulators complete
  const uint32_t bar_acc2 = smem_u32(&bars[2 * NB + 1]);
  const uint32_t ring_addr = smem_u32(smem + 128);

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2 * NB + 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init_fence();
  }
  for (int i = tid; i < 6 * C; i += THREADS) {
    const int cv = i / C, c = i - cv * C, j = cv >> 1;
    bias_s[i] = (j < nrb) ? ldg(((cv & 1) ? p.bias2[j] : p.bias1[j]) + c) : 0.f;
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  // TMEM allocations start at lane 0 and a column that is a multiple of 32 below 512: 4 votes rebuild it
  const uint32_t tmem_raw = *tmem_slot;
  const uint32_t tmem_base = uniform_bits(tmem_raw, 5, 9);
  if (tmem_base != tmem_raw) trap_now();

  const int n_ttiles = (T + 127) / 128;
  const int n_items = p.B * n_ttiles;
  const int my_items = (WETTS_BID < n_items) ? (n_items - WETTS_BID + WETTS_NBLK - 1) / WETTS_NBLK : 0;
  const long long bs = (long long)C * T;
  const float slope = p.slope, inv_slope = 1.0f / p.slope;

  // ---------------------------------------------------------------- activation prefetch / staging
  // unit u = (channel group cg = u % CG, row quad q = u / CG): four 16 B loads (4 channels x 4 consecutive
  // samples), transposed in registers into four (row, 4-channel) 16 B stores per hi/lo tile
  float4 pf[NU][4];
  auto prefetch = [&](int item, int j) {
    const int b = item / n_ttiles;
    const int t0 = (item - b * n_ttiles) * 128;
    const int H = (p.d1[j] + p.d2[j]) * (p.k[j] - 1) / 2;
    const int Hp = (H + 3) & ~3;
    const int Q = (128 + 2 * Hp) >> 2;
    const float* in_b = p.in + (long long)b * bs;
    if (SW0 > 0 && warp < SW0) return;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int u = (tid - 32 * SW0) + i * STHREADS;
      const int cg = u & (CG - 1), q = u >> LOG_CG;
      const int t = t0 - Hp + 4 * q;
      const bool ok = (q < Q) && (t >= 0) && (t < T);     // T % 4 == 0 (checked on the host): a quad is all in or all out
      const float* src = in_b + (long long)(4 * cg) * T + t;
#pragma unroll
      for (int e = 0; e < 4; ++e) pf[i][e] = ok ? ldg4(src + (long long)e * T) : float4{0.f, 0.f, 0.f, 0.f};
    }
  };
  auto lrelu = [&](float x) { return fmaxf(x, x * slope); };                  // 0 < slope < 1
  auto inv_lrelu = [&](float y) { return fminf(y, y * inv_slope); };
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
  auto stage = [&](int j) {
    const int H = (p.d1[j] + p.d2[j]) * (p.k[j] - 1) / 2;
    const int Hp = (H + 3) & ~3;
    const int Q = (128 + 2 * Hp) >> 2;
    if (SW0 > 0 && warp < SW0) return;
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const int u = (tid - 32 * SW0) + i * STHREADS;
      const int cg = u & (CG - 1), q = u >> LOG_CG;
      if (q < Q) {
        split_store(cg, 4 * q + 0, lrelu(pf[i][0].x), lrelu(pf[i][1].x), lrelu(pf[i][2].x), lrelu(pf[i][3].x));
        split_store(cg, 4 * q + 1, lrelu(pf[i][0].y), lrelu(pf[i][1].y), lrelu(pf[i][2].y), lrelu(pf[i][3].y));
        split_store(cg, 4 * q + 2, lrelu(pf[i][0].z), lrelu(pf[i][1].z), lrelu(pf[i][2].z), lrelu(pf[i][3].z));
        split_store(cg, 4 * q + 3, lrelu(pf[i][0].w), lrelu(pf[i][1].w), lrelu(pf[i][2].w), lrelu(pf[i][3].w));
      }
    }
  };
  // pre-activation value recovered from the staged hi/lo pair of (row r, channel group cg)
  auto staged_value = [&](int cg, int r, float* x) {
    const uint8_t* src = Ahi + ((size_t)cg * Rp + r) * 16;
    const float4 hi = *reinterpret_cast<const float4*>(src);
    const float4 lo = *reinterpret_cast<const float4*>(src + a_half);
    x[0] = inv_lrelu(hi.x + lo.x);
    x[1] = inv_lrelu(hi.y + lo.y);
    x[2] = inv_lrelu(hi.z + lo.z);
    x[3] = inv_lrelu(hi.w + lo.w);
  };

  long long prof[kFusedRbProfPhases];
  long long t_prev = 0;
  if (PROFILE) {
#pragma unroll
    for (int i = 0; i < kFusedRbProfPhases; ++i) prof[i] = 0;
  }
  // ---------------------------------------------------------------- weight ring
  // Chunk number g = it*nq + q (q = index within the item) is pure arithmetic on loop counters and kernel
  // parameters, so slot = g % NB, phase = g / NB and every descriptor derived from them stay in uniform
  // registers: the tcgen05.mma operands need no R2UR (the issue loop is otherwise ~90 cycles per MMA).
  const uint32_t nq = (uint32_t)p.nq;
  const uint64_t w_policy = l2_policy_evict_last();   // weights stay L2-resident under the activation stream
  // (slot, number of earlier uses of the slot) of chunk q of this CTA's it-th item.  NB = 4: bit fields of
  // the chunk number.  NB = 6: nq is a multiple of 6 (host-checked), so the slot depends on q only;
  // q / 6 = (q * 171) >> 10 holds for q < 500.
  auto ring_pos = [&](uint32_t it_, uint32_t q_, uint32_t& slot, uint32_t& use) {
    if (NB == 4) {
      const uint32_t g = it_ * nq + q_;
      slot = g & 3u;
      use = g >> 2;
    } else {
      const uint32_t qd = (q_ * 171u) >> 10;
      slot = q_ - 6u * qd;
      use = it_ * (uint32_t)p.nq_ring + qd;
    }
  };
  // producer (warp 1, warp-uniform): request chunk q_p of this CTA's it_p-th item into slot g % NB once the
  // MMAs that read the slot's previous occupant (chunk g - NB) have completed
  auto produce = [&](uint32_t it_p, uint32_t q_p) {
    if (q_p >= nq) { q_p -= nq; it_p += 1; }
    if (it_p >= (uint32_t)my_items) return;
    uint32_t slot, use;
    ring_pos(it_p, q_p, slot, use);
    if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);
    // no lane may still be inside the parity wait when the slot is handed back to the MMA warp: the barrier
    // could then complete a second phase and the late lane would wait for a parity that never returns
    warp_sync();
    if (elect_one()) {
      mbar_expect_tx(bar_full + 8 * slot, CHUNK_BYTES);
      bulk_g2s_hint(ring_addr + slot * CHUNK_BYTES, reinterpret_cast<const uint8_t*>(p.w) + (size_t)q_p * CHUNK_BYTES,
                    CHUNK_BYTES, bar_full + 8 * slot, w_policy);
    }
    warp_sync();
  };
  // while warp 0 multiplies chunks [qbase, qbase + n) the producer refills every slot the moment its chunk
  // completes: requests [qbase + NB, qbase + n + NB).  The last request waits for the conv's last chunk, so
  // the producer leaves the phase together with the accumulator barrier and the next conv starts with a
  // full ring (its first NB chunks travel during the SIMT phase in between).
  auto produce_range = [&](uint32_t it, uint32_t qbase, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) produce(it, qbase + i + (uint32_t)PD);
  };
  if (warp == 1)
    for (int i = 0; i < PD; ++i) produce(0u, (uint32_t)i);

  // Descriptor start addresses are built from p.smem_off (the CTA-local offset of the dynamic shared memory
  // base, probed once on the host side and verified here) plus compile-time offsets: values that come from
  // kernel parameters are uniform by construction, whereas an address derived from the `smem` pointer may be
  // kept in a vector register by ptxas and then costs an R2UR per tcgen05.mma operand.
  if (tid == 0 && (smem_u32(smem) & 0xFFFFFFu) != p.smem_off) trap_now();
  const uint32_t idesc_n = idesc_tf32_m128(N), idesc_2n = idesc_tf32_m128(2 * N);
  const uint64_t adesc0 = make_desc(p.smem_off + 128u + NB * CHUNK_BYTES + 6u * C * 4u, (uint32_t)Rp * 16u, 128u);
  const uint64_t bdesc0 = make_desc(p.smem_off + 128u, (uint32_t)(2 * N) * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  const uint32_t a_lo_delta = a_half >> 4;

  // One conv on the tensor pipe (warp 0): for every tap and 32-channel slice, multiply the weight chunk with
  // `nblk` 128-row blocks of the activation tile (block m starts at row row0 + m*row_step + tap*dil).
  auto run_conv = [&](uint32_t it, uint32_t qbase, int k, int dil, int nblk, int row0, int row_step, uint32_t d_col0,
                      uint32_t done_bar, bool fresh) {
    for (int tap = 0; tap < k; ++tap) {
      for (int kh = 0; kh < KH; ++kh) {
        const uint32_t q = qbase + (uint32_t)(tap * KH + kh);
        uint32_t slot, use;
        ring_pos(it, q, slot, use);
        const uint32_t par = use & 1u;
        long long tw = 0;
        if (PROFILE) tw = clock_now();
        mbar_wait(bar_full + 8 * slot, par);
        if (PROFILE) prof[14] += clock_now() - tw;
        warp_sync();   // same ABA guard as in produce(): all lanes have seen this phase before the slot can recycle
        const uint32_t b0 = blo0 + slot * (CHUNK_BYTES >> 4);
        const uint32_t first = (fresh && tap == 0 && kh == 0) ? 0u : 1u;   // 0: overwrite the accumulators
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
  if (PROFILE) t_prev = clock_now();
  auto mark = [&](int phase) {
    if (PROFILE) {
      const long long now = clock_now();
      prof[phase] += now - t_prev;
      t_prev = now;
    }
  };
  if (my_items > 0) prefetch(WETTS_BID, 0);

  for (int it = 0; it < my_items; ++it) {
    const int item = WETTS_BID + it * WETTS_NBLK;
    const int b = item / n_ttiles;
    const int t0 = (item - b * n_ttiles) * 128;
    const int c0 = 16 * grp;                 // this thread's 16 output channels in both epilogues
    // sum over resblocks of (bias2_j + x1_j): the convolution part of the MRF sum accumulates in TMEM
    float racc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) racc[i] = 0.f;

    for (int j = 0; j < nrb; ++j) {
      const int k = p.k[j], d1 = p.d1[j], d2 = p.d2[j];
      const int h1 = d1 * (k - 1) / 2, h2 = d2 * (k - 1) / 2, H = h1 + h2;
      const int Hp = (H + 3) & ~3, dl = Hp - H;
      const int next_item = (j + 1 < nrb) ? item : ((it + 1 < my_items) ? item + WETTS_NBLK : -1);
      const int next_j = (j + 1 < nrb) ? j + 1 : 0;

      mark(10);
      stage(j);
      fence_async_smem();
      mark(0);
      tc_fence_before();
      cta_sync();
      tc_fence_after();
      mark(1);

      // ---- conv1: x1 on rows [dl + h1, R - dl - h1).  The next tile's input is requested while the tensor pipe
      // works; the issuer and the producer serve the pipe first (its instruction queue is shallow and the weight
      // stream must not queue behind the activation burst).
      if (warp == 0) {
        run_conv((uint32_t)it, (uint32_t)p.qoff[2 * j], k, d1, 2, dl, 2 * h2, 0u, bar_acc1, true);
      } else if (warp == 1) {
        produce_range((uint32_t)it, (uint32_t)p.qoff[2 * j], (uint32_t)(k * KH));
      }
      if (PROFILE) prof[11 + (j < 2 ? j : 2)] += clock_now() - t_prev;
      if (next_item >= 0) prefetch(next_item, next_j);
      mark(2);
      mbar_wait(bar_acc1, rb_count & 1);
      tc_fence_after();
      mark(3);
      {
        const float* bias = bias_s + (2 * j) * C + c0;
        const int n_new = 2 * h2;                                    // rows only block 1 provides
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          if (mb == 1 && 32 * q4 + 31 < 128 - n_new) continue;        // warp-uniform: nothing new in this lane quarter
          const int r1 = dl + (mb ? n_new : 0) + row_i + h1;
          const bool active = (mb == 0) || (row_i >= 128 - n_new);
          const int t = t0 - Hp + r1;
          const bool inside = (t >= 0) && (t < T);
          const uint32_t ta = tmem_base + lane_sel + (uint32_t)(mb * 2 * N + c0);
          float v[16], vs[16];
          tmem_ld16_nowait(ta, v);
          tmem_ld16_nowait(ta + (uint32_t)N, vs);
          tmem_ld_wait();
          if (active) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const int cg = (c0 >> 2) + g4;
              float x[4], y[4];
              staged_value(cg, r1, x);
              const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * g4);
              const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x1 = inside ? ((v[4 * g4 + e] + vs[4 * g4 + e]) + bb[e]) + x[e] : 0.f;
                y[e] = lrelu(x1);
              }
              split_store(cg, r1, y[0], y[1], y[2], y[3]);
            }
          }
        }
      }
      fence_async_smem();
      mark(4);
      tc_fence_before();
      cta_sync();
      tc_fence_after();
      mark(5);

      // ---- conv2: the 128 outputs of this item, accumulated over the resblocks in TMEM
      if (warp == 0) {
        run_conv((uint32_t)it, (uint32_t)p.qoff[2 * j + 1], k, d2, 1, dl + h1, 0, (uint32_t)(4 * N), bar_acc2, j == 0);
      } else if (warp == 1) {
        produce_range((uint32_t)it, (uint32_t)p.qoff[2 * j + 1], (uint32_t)(k * KH));
      }
      mark(6);
      {
        // residual x1 + bias of this resblock, read while the tensor pipe works
        const float* bias = bias_s + (2 * j + 1) * C + c0;
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          float x[4];
          staged_value((c0 >> 2) + g4, Hp + row_i, x);
          const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * g4);
          racc[4 * g4 + 0] += bv.x + x[0];
          racc[4 * g4 + 1] += bv.y + x[1];
          racc[4 * g4 + 2] += bv.z + x[2];
          racc[4 * g4 + 3] += bv.w + x[3];
        }
      }
      mbar_wait(bar_acc2, rb_count & 1);
      tc_fence_after();
      mark(7);
      if (j == nrb - 1) {
        float v[16], vs[16];
        tmem_ld16_nowait(tmem_base + lane_sel + (uint32_t)(4 * N + c0), v);
        tmem_ld16_nowait(tmem_base + lane_sel + (uint32_t)(5 * N + c0), vs);
        tmem_ld_wait();
        const int t = t0 + row_i;
        if (t < T) {
          float* op = p.out + (long long)b * bs + (long long)c0 * T + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float sum = (v[i] + vs[i]) + racc[i];
            st_streaming(op + (long long)i * T, (nrb > 1) ? sum / p.div : sum);
          }
        }
      }
      rb_count += 1;
      mark(8);
      // every thread is done with the activation tile and with TMEM before either is overwritten
      tc_fence_before();
      cta_sync();
      tc_fence_after();
      mark(9);
    }
  }
  if (PROFILE && p.prof && (tid == 0 || tid == 32)) {
    for (int i = 0; i < kFusedRbProfPhases; ++i)
      p.prof[((size_t)WETTS_BID * 2 + (tid ? 1 : 0)) * kFusedRbProfPhases + i] = prof[i];
  }
  cta_sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace wetts
