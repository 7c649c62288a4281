// Fused HiFi-GAN MRF stage, f16-split tcgen05 version (sm_100a): ResBlock2 AND ResBlock1 generators.
//
//   out = (1/nrb) * sum_j rb_j(x)
//   ResBlock2 (decoders.py:205-214):  for d in (d0, d1):      x = conv_{k,d}(lrelu x) + x
//   ResBlock1 (decoders.py:157-170):  for d in (d0, d1, d2):  x = conv_{k,1}(lrelu(conv_{k,d}(lrelu x))) + x
//   MRF mean (decoders.py:72-76).
//
// One launch per generator stage.  Work item = (utterance b, ITEM output samples), ITEM = 128 or 256 (a 256-sample item
// amortises the halo: 5 instead of 6 M blocks per 256 samples of a ResBlock2 pair, 13 % fewer staged rows; below,
// "128" stands for ITEM where the text describes the item).  For resblock j with total halo
// H = sum of the conv halos, Hp = H rounded up to 4, dl = Hp - H, the CTA stages R = 128 + 2Hp input rows of lrelu(x)
// as the tcgen05 A operand (M = time rows, K = channels, no-swizzle K-major canonical layout):
//     element (row r, channel c) at tile + (c/8)*RP*16 + r*16 + (c%8)*2      (fp16; hi tile, then the lo' tile)
// so a conv tap is a row shift of the descriptor start address.  Row r <-> sample t0 - Hp + r.  RP is odd, which keeps
// the staging stores (lanes across channel groups) and the epilogue accesses (lanes across rows) bank-conflict free.
// The convs of a resblock run one after the other on the tile, each on the rows the later convs still need
// (n = 128 + 2*remaining halo rows, as one or two M = 128 blocks; the second block covers the LAST 128 rows):
//   * a conv whose result feeds the next conv writes lrelu(result) back as hi/lo' fp16 -- IN PLACE over lrelu(x) for
//     the residual convs (legal: every MMA of the conv has completed, acc barrier), into the second tile T for
//     ResBlock1's inner conv (x must survive for the residual); rows outside [0, T) are written as zeros (the
//     reference's zero padding of the next conv's input);
//   * the residual is recovered from the staged lrelu value (hi + lo' * 2^-11, inverse lrelu; deviation <= 2^-22);
//   * the last conv's result (+ bias + residual) is accumulated over the resblocks in registers; `out` is written once.
//
// fp32 accuracy on the f16 tensor pipe:  x ~ hi + lo' * 2^-11 with hi = f16(x), lo' = f16((x - hi) * 2^11): 22
// significand bits like the 3xTF32 split, but K = 16 channels per tcgen05.mma (kind::f16), i.e. half the MMAs and half
// the shared-memory operand bytes per MAC (these small-N MMAs are bound by the 128 B/clk operand read).  Two MMAs per
// k-step (tc_mma_f16_split2): A_hi x [B_hi | B_lo'] -> accumulator columns [0, 2N), A_lo' x B_hi -> columns [N, 2N);
// the epilogue adds columns [N, 2N) scaled by 2^-11.  Inputs must satisfy |x| < 65504 (saturating conversion).
//
// Weights stream through a shared-memory ring of (tap, 32-input-channel) chunks filled by cp.async.bulk (L2
// evict_last); the chunk sequence is identical for every item, so the ring never drains between items.  Warp 0 only
// issues MMAs, warp 1 refills ring slots the moment their chunk completes; the next tile's activations are prefetched
// into registers while the tensor pipe works.  Accumulators: 2 blocks x 2N TMEM columns, reused by every conv, so
// several CTAs fit per SM (C = 32: 128 columns, ~60 KB of shared memory) and one CTA's SIMT phases (staging,
// epilogues) overlap another CTA's MMAs.
//
// This file contains no PTX: everything hardware specific is in tc_prims.cuh, and the same source runs in the host
// CTA emulator (tests/emu).
#pragma once
#include "fused_mrf16_args.h"
#include "tc_prims.cuh"

namespace wetts {

template <int C, int THREADS, int MINB, int NB, int RP, bool TWO_TILES, bool PROFILE = false, int ITEM = 128>
WETTS_GLOBAL void WETTS_LAUNCH_BOUNDS(THREADS, MINB) fused_mrf16_kernel(const FusedMrfArgs p) {
  using namespace tc;
  static_assert(C == 32 || C == 64 || C == 128, "channel count");
  static_assert(NB == 4 || NB == 6, "ring size");
  static_assert((RP & 1) == 1 && RP >= 225, "odd row pitch");
  static_assert(ITEM == 128 || ITEM == 256 || ITEM == 384, "output samples per work item");
  constexpr int OBLK = ITEM / 128;                              // M blocks of the last conv of a resblock (its output rows)
  constexpr int MBLK = OBLK + 1;                                // M blocks of any other conv (output + remaining halo <= 128 rows more)
  constexpr int N = C;
  constexpr int KH = C / 32;
  constexpr uint32_t CHUNK_BYTES = 4u * 2u * N * 16u;          // [4 k-groups][hi | lo' : 2N rows][8 halfs]
  // MBLK accumulator blocks x [hi*hi | small terms], rounded up to a power of two (TMEM allocation granularity)
  constexpr uint32_t TMEM_COLS = (MBLK * 2u * N <= 128u) ? 128u : ((MBLK * 2u * N <= 256u) ? 256u : 512u);
  static_assert(MBLK * 2 * N <= 512, "accumulator blocks exceed TMEM");
  constexpr int CG8 = C / 8;                                    // 8-channel groups of the activation tile
  constexpr int LOG_CG8 = (CG8 == 4) ? 2 : (CG8 == 8 ? 3 : 4);
  constexpr int NWARP = THREADS / 32, GRPS = NWARP / 4;         // warps sharing a TMEM lane quarter split the columns
  static_assert(NWARP % 4 == 0 && (C / 16) % GRPS == 0, "warps must tile the 128 x C accumulator block");
  constexpr int SL = (C / 16) / GRPS;                           // 16-channel slices per thread
  constexpr int QMAX = RP / 4;                                  // row quads of the tile
  constexpr int UNITS = CG8 * QMAX;                             // staging units (8 channels x 4 rows)
  // warps 0 (MMA issuer) and 1 (weight producer) stay out of the staging when the other warps cover the tile
  constexpr int SW0 = (UNITS <= THREADS - 64) ? 2 : 0;
  constexpr int STHREADS = THREADS - 32 * SW0;
  // one staging unit per thread: its 8 loads are prefetched into registers during the previous MMA phase; more units
  // than threads (C = 64 with 256 threads): NU units per thread, loaded and stored inside the staging phase
  // (256-sample items: two units per thread are prefetched)
  constexpr int NU = (UNITS + STHREADS - 1) / STHREADS;
  constexpr bool PREFETCH = (NU == 1) || (ITEM > 128 && NU == 2);
  constexpr int NPF = PREFETCH ? NU : 1;
  constexpr uint32_t A_HALF = (uint32_t)CG8 * RP * 16u;         // bytes of the hi (or lo') tile
  constexpr uint32_t TILE_BYTES = 2u * A_HALF;
  constexpr int NBIAS = kMrfMaxRb * kMrfMaxConv;

  WETTS_SMEM_DECL(smem);
  const int tid = WETTS_TID, lane = tid & 31;
  const int warp = (int)uniform_bits((uint32_t)(tid >> 5), 0, 5);
  const int T = p.T, nrb = p.nrb, nconv = p.nconv;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 120);
  float* bias_s = reinterpret_cast<float*>(smem + 128 + NB * CHUNK_BYTES);     // [nrb][nconv][C]
  uint8_t* Xt = smem + 128 + NB * CHUNK_BYTES + NBIAS * C * 4;
  uint8_t* Tt = Xt + (TWO_TILES ? TILE_BYTES : 0);
  const uint32_t bar_full = smem_u32(&bars[0]);        // [NB] TMA -> MMA: weight chunk landed
  const uint32_t bar_empty = smem_u32(&bars[NB]);      // [NB] MMA -> TMA: weight slot reusable
  const uint32_t bar_acc = smem_u32(&bars[2 * NB]);    //      accumulators of the current conv complete
  const uint32_t ring_addr = smem_u32(smem + 128);

  if ((smem_u32(smem) & 0xFFFFFFu) != p.smem_off) trap_now();     // same value in every thread: a uniform branch
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  if (tid == 0) {
    for (int i = 0; i < 2 * NB + 1; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init_fence();
  }
  for (int i = tid; i < NBIAS * C; i += THREADS) {
    const int cv = i / C, c = i - cv * C, j = cv / kMrfMaxConv, cc = cv - j * kMrfMaxConv;
    bias_s[i] = (j < nrb && cc < nconv) ? ldg(p.bias[j][cc] + c) : 0.f;
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  // TMEM allocations start at lane 0 and a column that is a multiple of 32 below 512: 4 votes rebuild it
  const uint32_t tmem_raw = *tmem_slot;
  const uint32_t tmem_base = uniform_bits(tmem_raw, 5, 9);
  if (tmem_base != tmem_raw) trap_now();

  const int n_ttiles = (T + ITEM - 1) / ITEM;
  const int n_items = p.item_map ? ldg_i32(p.n_items_dev) : p.B * n_ttiles;
  const int my_items = (WETTS_BID < n_items) ? (n_items - WETTS_BID + WETTS_NBLK - 1) / WETTS_NBLK : 0;
  const long long bs = (long long)C * T;
  const float slope = p.slope, inv_slope = 1.0f / p.slope;

  // item -> (utterance, first sample).  No data-dependent control flow here: a loop whose trip count depends on
  // loaded data would make everything after it "divergent" for the compiler (and the MMA operands go through R2UR).
  auto decode = [&](int item, int& b, int& t0) {
    if (p.item_map) {
      b = ldg_i32(&p.item_map[item].x);
      t0 = ldg_i32(&p.item_map[item].y);
    } else {
      b = item / n_ttiles;
      t0 = (item - b * n_ttiles) * ITEM;
    }
  };
  auto halo_of = [&](int j) {
    int H = 0;
    for (int c = 0; c < nconv; ++c) H += p.dil[j][c] * (p.k[j] - 1) / 2;
    return H;
  };

  // ---------------------------------------------------------------- activation prefetch / staging
  // unit u = (8-channel group cg = u % CG8, row quad q = u / CG8): eight 16 B loads (8 channels x 4 consecutive
  // samples), transposed in registers into four (row, 8-channel) 16 B stores per hi / lo' tile
  float4 pf[NPF][8];
  const int su = tid - 32 * SW0;                       // staging unit of this thread (negative: not a staging thread)
  auto prefetch = [&](int item, int j) {
    if (!PREFETCH) return;
    if (su < 0 || su >= UNITS) return;
    int b, t0;
    decode(item, b, t0);
    const int Hp = (halo_of(j) + 3) & ~3;
    const int Q = (ITEM + 2 * Hp) >> 2;
#pragma unroll
    for (int i = 0; i < NPF; ++i) {
      const int u = su + i * STHREADS;
      const int cg = u & (CG8 - 1), q = u >> LOG_CG8;
      const int t = t0 - Hp + 4 * q;
      const bool ok = (u < UNITS) && (q < Q) && (t >= 0) && (t < T);     // T % 4 == 0 (host-checked): a quad is all in or all out
      const float* src = p.in + (long long)b * bs + (long long)(8 * cg) * T + t;
#pragma unroll
      for (int e = 0; e < 8; ++e) pf[i][e] = ok ? ldg4(src + (long long)e * T) : float4{0.f, 0.f, 0.f, 0.f};
    }
  };
  auto lrelu = [&](float x) { return fmaxf(x, x * slope); };                  // 0 < slope < 1
  auto inv_lrelu = [&](float y) { return fminf(y, y * inv_slope); };
  // 8 channels of one row -> hi / lo' fp16, one 16 B store per tile half
  auto split_store8 = [&](uint8_t* tile, int cg, int r, const float* y) {
    uint4 hi, lo;
    f16_split2(y[0], y[1], hi.x, lo.x);
    f16_split2(y[2], y[3], hi.y, lo.y);
    f16_split2(y[4], y[5], hi.z, lo.z);
    f16_split2(y[6], y[7], hi.w, lo.w);
    uint8_t* dst = tile + ((size_t)cg * RP + r) * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + A_HALF) = lo;
  };
  auto store_unit = [&](const float4 (&pu)[8], int cg, int q) {
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = lrelu(pu[e].x);
    split_store8(Xt, cg, 4 * q + 0, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = lrelu(pu[e].y);
    split_store8(Xt, cg, 4 * q + 1, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = lrelu(pu[e].z);
    split_store8(Xt, cg, 4 * q + 2, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) y[e] = lrelu(pu[e].w);
    split_store8(Xt, cg, 4 * q + 3, y);
  };
  auto stage = [&](int item, int j) {
    if (su < 0) return;
    const int Hp = (halo_of(j) + 3) & ~3;
    const int Q = (ITEM + 2 * Hp) >> 2;
    if (PREFETCH) {
      if (su >= UNITS) return;
#pragma unroll
      for (int i = 0; i < NPF; ++i) {
        const int u = su + i * STHREADS;
        const int cg = u & (CG8 - 1), q = u >> LOG_CG8;
        if (u < UNITS && q < Q) store_unit(pf[i], cg, q);
      }
    } else {
      int b, t0;
      decode(item, b, t0);
#pragma unroll 1
      for (int i = 0; i < NU; ++i) {
        const int u = su + i * STHREADS;
        const int cg = u & (CG8 - 1), q = u >> LOG_CG8;
        if (u < UNITS && q < Q) {
          const int t = t0 - Hp + 4 * q;
          const bool ok = (t >= 0) && (t < T);
          const float* src = p.in + (long long)b * bs + (long long)(8 * cg) * T + t;
#pragma unroll
          for (int e = 0; e < 8; ++e) pf[0][e] = ok ? ldg4(src + (long long)e * T) : float4{0.f, 0.f, 0.f, 0.f};
          store_unit(pf[0], cg, q);
        }
      }
    }
  };
  // pre-activation values of 8 channels recovered from the staged hi / lo' pair of (row r, channel group cg)
  auto staged_value8 = [&](const uint8_t* tile, int cg, int r, float* x) {
    const uint8_t* src = tile + ((size_t)cg * RP + r) * 16;
    const uint4 hi = *reinterpret_cast<const uint4*>(src);
    const uint4 lo = *reinterpret_cast<const uint4*>(src + A_HALF);
    f16_join2(hi.x, lo.x, x[0], x[1]);
    f16_join2(hi.y, lo.y, x[2], x[3]);
    f16_join2(hi.z, lo.z, x[4], x[5]);
    f16_join2(hi.w, lo.w, x[6], x[7]);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = inv_lrelu(x[e]);
  };

  long long prof[kMrfProfPhases];
  long long t_prev = 0;
  if (PROFILE) {
#pragma unroll
    for (int i = 0; i < kMrfProfPhases; ++i) prof[i] = 0;
  }
  // ---------------------------------------------------------------- weight ring
  // Chunk number g = it*nq + q (q = index within the item) is pure arithmetic on loop counters and kernel
  // parameters, so slot, phase and every descriptor derived from them stay in uniform registers.
  const uint32_t nq = (uint32_t)p.nq;
  const uint64_t w_policy = l2_policy_evict_last();   // weights stay L2-resident under the activation stream
  auto ring_pos = [&](uint32_t it_, uint32_t q_, uint32_t& slot, uint32_t& use) {
    if (NB == 4) {
      const uint32_t g = it_ * nq + q_;
      slot = g & 3u;
      use = g >> 2;
    } else {                               // nq % 6 == 0 (host-checked): the slot depends on q only
      const uint32_t qd = (q_ * 171u) >> 10;           // q / 6 for q < 500
      slot = q_ - 6u * qd;
      use = it_ * (uint32_t)p.nq_ring + qd;
    }
  };
  // producer (warp 1, warp-uniform): request chunk q_p of this CTA's it_p-th item once the MMAs that read the slot's
  // previous occupant have completed
  auto produce = [&](uint32_t it_p, uint32_t q_p) {
    if (q_p >= nq) { q_p -= nq; it_p += 1; }
    if (it_p >= (uint32_t)my_items) return;
    uint32_t slot, use;
    ring_pos(it_p, q_p, slot, use);
    if (use > 0) mbar_wait(bar_empty + 8 * slot, (use - 1) & 1);
    // no lane may still be inside the parity wait when the slot is handed back to the MMA warp (ABA guard)
    warp_sync();
    if (elect_one()) {
      mbar_expect_tx(bar_full + 8 * slot, CHUNK_BYTES);
      bulk_g2s_hint(ring_addr + slot * CHUNK_BYTES, reinterpret_cast<const uint8_t*>(p.w) + (size_t)q_p * CHUNK_BYTES,
                    CHUNK_BYTES, bar_full + 8 * slot, w_policy);
    }
    warp_sync();
  };
  auto produce_range = [&](uint32_t it, uint32_t qbase, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) produce(it, qbase + i + (uint32_t)NB);
  };
  if (warp == 1)
    for (int i = 0; i < NB; ++i) produce(0u, (uint32_t)i);

  // descriptors from p.smem_off (kernel parameter: uniform by construction) plus compile-time offsets; the base is
  // verified at kernel entry (a thread-dependent branch to a trap HERE makes ptxas treat the issue loop below as
  // possibly diverged: every UTCHMMA predicated, operands through R2UR)
  const uint32_t idesc_n = idesc_f16_m128(N), idesc_2n = idesc_f16_m128(2 * N);
  constexpr uint32_t X_OFF = 128u + NB * CHUNK_BYTES + (uint32_t)NBIAS * C * 4u;
  const uint64_t adesc0 = make_desc(p.smem_off + X_OFF, (uint32_t)RP * 16u, 128u);
  const uint64_t bdesc0 = make_desc(p.smem_off + 128u, (uint32_t)(2 * N) * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  constexpr uint32_t A_LO_DELTA = A_HALF >> 4, T_DELTA = TILE_BYTES >> 4;

  // One conv on the tensor pipe (warp 0): for every tap and 32-channel slice, multiply the weight chunk with `nblk`
  // 128-row blocks of the source tile (block m starts at row row0 + m*128 + tap*dil; the last block at row0 + n_out - 128,
  // so it ends with the conv's last output row and may overlap its predecessor).
  // ring_base = low bits of this item's first chunk number: only slot and phase parity matter to the consumer.
  // Issue-path hygiene (SASS-verified: unpredicated UTCHMMA fed by UIADD3/UMOV only): all scalars here derive from kernel
  // parameters and loop counters; the epilogue loops use structured ifs only (no `continue` under a thread-dependent
  // condition); no thread-dependent branch to a trap after the prologue.
  auto run_conv = [&](uint32_t ring_base, uint32_t qbase, int k, int dil, int nblk, int row0, int n_out, uint32_t tile_delta) {
    for (int tap = 0; tap < k; ++tap) {
      for (int kh = 0; kh < KH; ++kh) {
        const uint32_t q = qbase + (uint32_t)(tap * KH + kh);
        uint32_t slot, par;
        if (NB == 4) {
          const uint32_t g = (ring_base + q) & 7u;
          slot = g & 3u;
          par = g >> 2;
        } else {
          const uint32_t qd = (q * 171u) >> 10;          // q / 6 for q < 500
          slot = q - 6u * qd;
          par = (ring_base + qd) & 1u;
        }
        long long tw = 0;
        if (PROFILE) tw = clock_now();
        mbar_wait(bar_full + 8 * slot, par);
        if (PROFILE) prof[7] += clock_now() - tw;
        warp_sync();   // ABA guard: all lanes have seen this phase before the slot can recycle
        const uint32_t b0 = blo0 + slot * (CHUNK_BYTES >> 4);
        const uint32_t first = (tap == 0 && kh == 0) ? 0u : 1u;   // 0: overwrite the accumulators
        for (int m = 0; m < nblk; ++m) {
          // (ITEM == 128: at most two blocks, written as in the 128-sample kernel this generalises -- same code as before)
          const int m_row = (ITEM == 128) ? m * (n_out - 128) : ((m == nblk - 1) ? n_out - 128 : m * 128);
          const uint32_t a0 = alo0 + tile_delta + (uint32_t)((kh * 4) * RP + row0 + m_row + tap * dil);
          const uint32_t d_tmem = tmem_base + (uint32_t)(m * 2 * N);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t al = a0 + (uint32_t)(kk * 2 * RP), bl = b0 + (uint32_t)(kk * 2 * 2 * N);
            tc_mma_f16_split2(d_tmem, d_tmem + (uint32_t)N, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + A_LO_DELTA),
                              desc_with_lo(bdesc0, bl), idesc_2n, idesc_n, (kk == 0) ? first : 1u);
          }
        }
        if (elect_one()) tc_commit(bar_empty + 8 * slot);
        warp_sync();
      }
    }
    if (elect_one()) tc_commit(bar_acc);
    warp_sync();
  };

  // ---------------------------------------------------------------- main loop
  const int q4 = warp & 3, grp = warp >> 2;
  const int row_i = 32 * q4 + lane;                   // TMEM lane = row of the 128-row block
  const uint32_t lane_sel = (uint32_t)(32 * q4) << 16;
  uint32_t conv_count = 0;
  if (PROFILE) t_prev = clock_now();
  auto mark = [&](int phase) {
    if (PROFILE) {
      const long long now = clock_now();
      prof[phase] += now - t_prev;
      t_prev = now;
    }
  };
  if (p.stagger > 0) spin_cycles((long long)(WETTS_BID / p.n_sm) * p.stagger);
  if (my_items > 0) prefetch(WETTS_BID, 0);

  for (int it = 0; it < my_items; ++it) {
    const int item = WETTS_BID + it * WETTS_NBLK;
    int b, t0;
    decode(item, b, t0);
    float racc[OBLK][SL][16];
#pragma unroll
    for (int o = 0; o < OBLK; ++o)
#pragma unroll
      for (int s = 0; s < SL; ++s)
#pragma unroll
        for (int i = 0; i < 16; ++i) racc[o][s][i] = 0.f;

    for (int j = 0; j < nrb; ++j) {
      const int k = p.k[j];
      const int H = halo_of(j), Hp = (H + 3) & ~3;
      const int R = ITEM + 2 * Hp;
      const int next_item = (j + 1 < nrb) ? item : ((it + 1 < my_items) ? item + WETTS_NBLK : -1);
      const int next_j = (j + 1 < nrb) ? j + 1 : 0;

      mark(8);
      stage(item, j);
      fence_async_smem();
      mark(0);
      tc_fence_before();
      cta_sync();
      tc_fence_after();
      mark(1);

      int lo = Hp - H;                                 // first valid row of the source tile
      for (int c = 0; c < nconv; ++c) {
        const int d = p.dil[j][c], h = d * (k - 1) / 2;
        const int lo_out = lo + h, n_out = R - 2 * lo_out;
        const int nblk = (ITEM == 128) ? ((n_out > 128) ? 2 : 1) : ((n_out + 127) >> 7);   // <= MBLK (host-checked); the last conv: exactly OBLK
        const int ov = 128 * nblk - n_out;               // rows the last block shares with its predecessor
        const bool last = (c == nconv - 1);
        const bool inner = TWO_TILES && ((c & 1) == 0);      // ResBlock1's first conv of a pair: X -> T, no residual
        const bool from_t = TWO_TILES && ((c & 1) == 1);
        if (warp == 0) {
          const uint32_t ring_base = (NB == 4) ? (((uint32_t)it * nq) & 7u) : (((uint32_t)it * (uint32_t)p.nq_ring) & 1u);
          run_conv(ring_base, (uint32_t)p.qoff[j][c], k, d, nblk, lo, n_out, from_t ? T_DELTA : 0u);
        } else if (warp == 1) {
          produce_range((uint32_t)it, (uint32_t)p.qoff[j][c], (uint32_t)(k * KH));
        }
        // the next tile's input is requested while the tensor pipe works on the first conv of this resblock
        if (c == 0 && next_item >= 0) prefetch(next_item, next_j);
        mark(2);
        mbar_wait(bar_acc, conv_count & 1);
        conv_count += 1;
        tc_fence_after();
        mark(3);
        const float* bias = bias_s + (j * kMrfMaxConv + c) * C;
        uint8_t* dst_tile = inner ? Tt : Xt;
        // NOTE: structured ifs only (no `continue` under a thread-dependent condition): an induction variable that
        // joins at a divergent branch makes the loop exit -- and every MMA operand after it -- "divergent" for ptxas.
#pragma unroll 1
        for (int mb = 0; mb < nblk; ++mb) {
          const bool tail = (ITEM == 128) ? (mb != 0) : ((mb > 0) && (mb == nblk - 1));   // the block that may overlap
          const bool quarter_has_rows = !tail || (32 * q4 + 31 >= ov);                  // warp-uniform
          if (quarter_has_rows) {
            const bool active = !tail || (row_i >= ov);
            const int r = lo_out + (tail ? n_out - 128 : 128 * mb) + row_i;
            const int t = t0 - Hp + r;
            const bool inside = (t >= 0) && (t < T);
#pragma unroll
            for (int s = 0; s < SL; ++s) {
              const int c0 = 16 * (grp * SL + s);
              const uint32_t ta = tmem_base + lane_sel + (uint32_t)(mb * 2 * N + c0);
              float v[16], vs[16];
              tmem_ld16_nowait(ta, v);
              tmem_ld16_nowait(ta + (uint32_t)N, vs);
              tmem_ld_wait();
              if (active) {
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  const int cg = (c0 >> 3) + g8;
                  float val[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) val[e] = (v[8 * g8 + e] + vs[8 * g8 + e] * kF16LoInv) + bias[c0 + 8 * g8 + e];
                  if (!inner) {
                    float x[8];
                    staged_value8(Xt, cg, r, x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) val[e] += x[e];
                  }
                  if (last) {
                    // (the last conv has exactly OBLK blocks, none overlapping: mb indexes the register accumulator)
#pragma unroll
                    for (int o = 0; o < OBLK; ++o)
                      if (OBLK == 1 || o == mb) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) racc[o][s][8 * g8 + e] += val[e];
                      }
                  } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) val[e] = inside ? lrelu(val[e]) : 0.f;
                    split_store8(dst_tile, cg, r, val);
                  }
                }
              }
            }
          }
        }
        if (!last) fence_async_smem();
        mark(4);
        // every thread is done with TMEM (and with the tile it read) before the next conv / the next staging
        tc_fence_before();
        cta_sync();
        tc_fence_after();
        mark(5);
        lo = lo_out;
      }
    }
    // ---- MRF mean of this item
#pragma unroll
    for (int o = 0; o < OBLK; ++o) {
      const int t = t0 + 128 * o + row_i;
      if (t < T) {
#pragma unroll
        for (int s = 0; s < SL; ++s) {
          float* op = p.out + (long long)b * bs + (long long)(16 * (grp * SL + s)) * T + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) st_streaming(op + (long long)i * T, (nrb > 1) ? racc[o][s][i] / p.div : racc[o][s][i]);
        }
      }
    }
    mark(6);
  }
  if (PROFILE && p.prof && (tid == 0 || tid == 32)) {
    for (int i = 0; i < kMrfProfPhases; ++i)
      p.prof[((size_t)WETTS_BID * 2 + (tid ? 1 : 0)) * kMrfProfPhases + i] = prof[i];
  }
  cta_sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace wetts
