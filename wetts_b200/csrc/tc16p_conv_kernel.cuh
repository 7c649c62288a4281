// Pipelined per-layer implicit-GEMM Conv1d on the tensor pipe (sm_100a tcgen05, f16 operand split).
//
// Same arithmetic, tiling and packed-weight layout as conv1d_tc16_kernel (tc16_conv_kernel.cu), different pipeline:
// in that kernel every warp stages, then every warp runs the epilogue, with a CTA-wide barrier per work item, so
// the tensor pipe waits for the staging latency at the start of every item and idles through the epilogue (measured:
// 67 k cycles per flow in_layer item for 23 k cycles of MMAs).  Here the roles never meet at a CTA barrier:
//
//   warp 0          MMA issuer      wait acc_empty, then per chunk: wait b_full, per tile: wait a_full -> MMAs -> commit
//                                    a_free; commit b_free; at the end of the item commit acc_full
//   warp 1          weight producer  per chunk: wait b_free (NB-slot ring, 2..4) -> cp.async.bulk -> b_full
//   workers         stagers          per (chunk, tile): wait a_free (NA-slot ring) -> global loads, lrelu / masks / zero
//                                    padding, fp32 -> f16 split -> a_full
//                   epilogue         wait acc_full -> TMEM -> fused epilogue (tc_epilogue.cuh) -> acc_empty
//
// Two assignments of the 14 worker warps (TcConvArgs::all_warps):
//   0 "split":      warps 2..7 stage, warps 8..15 run the epilogue; the stagers run ahead into the next item while the
//                   epilogue warps drain the accumulators.  With one accumulator set in TMEM (N = 128, two M blocks =
//                   512 columns) the MMAs cannot start before the drain ends, so the epilogue warps idle through the MMAs
//                   and the stagers (192 threads, one load round trip per 16 rows x channels each) are the slow stage:
//                   measured slower than conv1d_tc16_kernel (38.5 vs 30.0 ms per step).
//   1 "all warps":  warps 2..15 stage the item (448 threads, two units = 32 loads in flight per thread), then warps 4..15
//                   drain it (three warps per TMEM lane quarter, 16-column slices round-robin).  Per item: max(MMA,
//                   staging) + epilogue, with the MMA / weight pipeline of this kernel (measured floor with staging and
//                   epilogue work skipped: 29.0 ms vs 38.8 ms for the v1 generator's 41 launches).
//
// Every role walks the same item sequence with its own counters.  Zero-tile items of the length-aware mode run no
// chunks; only the accumulator hand-shake ticks.
//
// No PTX here (tc_prims.cuh wrappers): the same source runs in the host CTA emulator (tests/emu/tc16p_emu.cpp).
#pragma once
#include "conv_args.h"
#include "tc_epilogue.cuh"
#include "tc_prims.cuh"

namespace wetts {

constexpr int kTc16pThreads = 512;
constexpr int kTc16pNA = 4;            // activation ring slots (maximum; 2 when 4 do not fit in shared memory)
constexpr int kTc16pWorkerWarp0 = 2;   // warps 2..15 are workers; see the role table above

constexpr int kTc16pNB = 4;            // weight ring slots (maximum)
// shared memory: [bars 192 B][A ring: NA x (hi | lo')][B ring: NB x weight tile].  The weight stream is what the MMAs
// wait for (a 41 KB chunk of a flow in_layer feeds 1.9 k cycles of MMAs, less than one bulk-copy round trip to L2), so
// the launcher deepens the weight ring first and the activation ring with what is left.
inline size_t tc16p_smem_bytes(int K, int dil, int N, int KC, int MB, int na = kTc16pNA, int nb = 2) {
  const int R = 128 * MB + (K - 1) * dil;
  const int Rp = (R + 7) & ~7;
  return 192 + (size_t)na * (2 * (size_t)(KC / 8) * Rp * 16) + (size_t)nb * ((size_t)K * (KC / 8) * 2 * N * 16);
}

// PROF: clock64 phase counters of one thread per role (WETTS_TC16P_PROFILE=1; table printed by the launcher)
constexpr int kTc16pProfSlots = 8;
template <bool PROF>
WETTS_GLOBAL void WETTS_LAUNCH_BOUNDS(kTc16pThreads, 1) conv1d_tc16p_kernel(const TcConvArgs p) {
  using namespace tc;
  long long pc[kTc16pProfSlots];
#pragma unroll
  for (int i = 0; i < kTc16pProfSlots; ++i) pc[i] = 0;
  long long t_a = 0, t_b = 0;
  const long long t_start = PROF ? clock_now() : 0;
  constexpr int NAMAX = kTc16pNA;
  WETTS_SMEM_DECL(smem);
  const ConvArgs& a = p.c;
  const int tid = WETTS_TID, lane = tid & 31;
  const int warp = (int)uniform_bits((uint32_t)(tid >> 5), 0, 4);
  const int K = a.K, dil = a.dil, T = a.T;
  const int N = p.N, KC = p.KC, MB = p.MB, MT = 128 * p.MB;
  const int R = MT + (K - 1) * dil;
  const int Rp = p.R_pad;
  const uint32_t a_half = (uint32_t)(KC / 8) * Rp * 16;
  const uint32_t a_bytes = 2 * a_half;
  const uint32_t b_bytes = (uint32_t)K * (KC / 8) * 2 * N * 16;
  const uint32_t tmem_cols = (uint32_t)p.tmem_cols;
  const bool aw = p.all_warps != 0;
  const int n_stagers = aw ? 14 * 32 : 6 * 32;                  // threads that stage (warps 2..15 / 2..7)
  const int epi_warp0 = aw ? 4 : 8, n_epi_warps = aw ? 12 : 8;  // warps that drain TMEM (.. 15)
  const int nparts = n_epi_warps / 4;                           // epilogue warps per TMEM lane quarter
  const uint32_t NA = (uint32_t)p.n_abuf, na_log = (p.n_abuf == 4) ? 2u : 1u;      // ring of 2 or 4 slots
  const uint32_t NB = (uint32_t)p.n_bbuf;                                           // 2, 3 or 4 slots

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 184);
  uint8_t* A0 = smem + 192;
  uint8_t* B0 = A0 + (size_t)NA * a_bytes;
  const uint32_t bar_a_free = smem_u32(&bars[0]);    // [NA] MMA -> stagers
  const uint32_t bar_a_full = smem_u32(&bars[4]);    // [NA] stagers -> MMA
  const uint32_t bar_b_full = smem_u32(&bars[8]);    // [4]  bulk copy -> MMA
  const uint32_t bar_b_free = smem_u32(&bars[12]);   // [4]  MMA -> producer
  const uint32_t bar_acc_full = smem_u32(&bars[16]); // [2]  MMA -> epilogue warps
  const uint32_t bar_acc_empty = smem_u32(&bars[18]);// [2]  epilogue warps -> MMA
  // accumulator slots in TMEM: 2 when two items' accumulators fit (N <= 64 with two M blocks): the MMAs of item i + 1
  // then run while item i is drained; slot = item count & 1
  const uint32_t S = (uint32_t)p.acc_slots, s_log = S - 1u;
  const uint32_t slot_cols = (uint32_t)(p.G * MB * 2 * N);
  const uint32_t A_addr = smem_u32(A0), B_addr = smem_u32(B0);

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < NAMAX; ++i) {
      mbar_init(bar_a_free + 8 * i, 1);
      mbar_init(bar_a_full + 8 * i, (uint32_t)(n_stagers / 32));
    }
    for (int i = 0; i < kTc16pNB; ++i) {
      mbar_init(bar_b_full + 8 * i, 1);
      mbar_init(bar_b_free + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, (uint32_t)n_epi_warps);
    }
    mbar_init_fence();
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = uniform_bits(*tmem_slot, 5, 9);

  const int G = p.G;
  const int group_rows = G * MT;
  const int n_groups = (T + group_rows - 1) / group_rows;
  const int items_per_nt = a.B * n_groups;
  const int n_items = items_per_nt * p.n_tiles;
  const int nbulk = (int)((b_bytes + 32767u) / 32768u);

  // item -> (n tile, utterance, first row, tiles, chunks): identical arithmetic in every role
  auto decode = [&](int item, int& nt, int& b, int& t_group0, int& tiles, int& nch) {
    int rem;
    if (p.nt_minor) {      // the N tiles of one row block in consecutive items (neighbouring CTAs): their reads meet in L2
      rem = item / p.n_tiles;
      nt = item - rem * p.n_tiles;
    } else {
      nt = item / items_per_nt;
      rem = item - nt * items_per_nt;
    }
    b = rem / n_groups;
    t_group0 = (rem - b * n_groups) * group_rows;
    tiles = (T - t_group0 + MT - 1) / MT;
    tiles = tiles < G ? tiles : G;
    if (a.la_len) tiles = ((long long)t_group0 >= (ldg_i64(a.la_len + b) + a.la_margin) * (long long)a.la_rate) ? 0 : tiles;
    nch = (tiles > 0) ? p.n_chunks : 0;
  };

  if (warp == 0) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc_n = idesc_f16_m128(N), idesc_2n = idesc_f16_m128(2 * N);
    const uint32_t a_lo_delta = a_half >> 4;
    uint32_t a_cnt = 0, bb = 0, b_use = 0, it_cnt = 0;   // weight ring position: slot bb, earlier uses of it b_use
    for (int item = WETTS_BID; item < n_items; item += WETTS_NBLK) {
      int nt, b, t_group0, tiles, nch;
      decode(item, nt, b, t_group0, tiles, nch);
      const uint32_t slot = it_cnt & s_log, s_use = it_cnt >> s_log;
      if (PROF) t_a = clock_now();
      if (s_use > 0) mbar_wait(bar_acc_empty + 8 * slot, (s_use - 1) & 1);   // the slot's previous item is drained
      if (PROF) pc[0] += clock_now() - t_a;
      tc_fence_after();
      for (int c = 0; c < nch; ++c) {
        if (PROF) t_a = clock_now();
        mbar_wait(bar_b_full + 8 * bb, b_use & 1);
        if (PROF) pc[1] += clock_now() - t_a;
        for (int g = 0; g < tiles; ++g) {
          const uint32_t ab = a_cnt & (NA - 1u);
          if (PROF) t_a = clock_now();
          mbar_wait(bar_a_full + 8 * ab, (a_cnt >> na_log) & 1);
          if (PROF) { t_b = clock_now(); pc[2] += t_b - t_a; }
          tc_fence_after();
          const uint64_t adesc0 = make_desc(A_addr + ab * a_bytes, (uint32_t)Rp * 16, 128);
          const uint64_t bdesc0 = make_desc(B_addr + bb * b_bytes, (uint32_t)(2 * N) * 16, 128);
          const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t d_tmem = tmem_base + slot * slot_cols + (uint32_t)((g * MB + mb) * 2 * N);
            for (int tap = 0; tap < K; ++tap) {
              uint32_t al = alo0 + (uint32_t)(mb * 128 + tap * dil);
              uint32_t bl = blo0 + (uint32_t)tap * (uint32_t)((KC / 8) * 2 * N);
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint32_t first = (c == 0 && tap == 0 && kk == 0) ? 0u : 1u;
                tc_mma_f16_split2(d_tmem, d_tmem + (uint32_t)N, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + a_lo_delta),
                                  desc_with_lo(bdesc0, bl), idesc_2n, idesc_n, first);
                al += 2u * (uint32_t)Rp;
                bl += 2u * (uint32_t)(2 * N);
              }
            }
          }
          if (elect_one()) tc_commit(bar_a_free + 8 * ab);
          warp_sync();
          if (PROF) pc[3] += clock_now() - t_b;
          a_cnt += 1;
        }
        if (elect_one()) tc_commit(bar_b_free + 8 * bb);
        warp_sync();
        bb += 1;
        if (bb == NB) { bb = 0; b_use += 1; }
      }
      if (elect_one()) tc_commit(bar_acc_full + 8 * slot);
      warp_sync();
      it_cnt += 1;
    }
  } else if (warp == 1) {
    // =============================== weight producer ===============================
    uint32_t bb = 0, use = 0;
    for (int item = WETTS_BID; item < n_items; item += WETTS_NBLK) {
      int nt, b, t_group0, tiles, nch;
      decode(item, nt, b, t_group0, tiles, nch);
      for (int c = 0; c < nch; ++c) {
        if (PROF) t_a = clock_now();
        if (use > 0) mbar_wait(bar_b_free + 8 * bb, (use - 1) & 1);
        if (PROF) pc[0] += clock_now() - t_a;
        warp_sync();
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wtc) + ((size_t)nt * p.n_chunks + c) * b_bytes;
        if (elect_one()) {
          mbar_expect_tx(bar_b_full + 8 * bb, b_bytes);
          for (int q = 0; q < nbulk; ++q) {
            const uint32_t off = (uint32_t)q * 32768u;
            const uint32_t n = (b_bytes - off) < 32768u ? (b_bytes - off) : 32768u;
            bulk_g2s(B_addr + bb * b_bytes + off, src + off, n, bar_b_full + 8 * bb);
          }
        }
        warp_sync();
        bb += 1;
        if (bb == NB) { bb = 0; use += 1; }
      }
    }
  } else {
    // =============================== workers: stage, then / or drain ===============================
    const bool do_stage = aw || warp < 8;
    const bool do_epi = warp >= epi_warp0;
    const int st = tid - 32 * kTc16pWorkerWarp0;          // 0 .. n_stagers - 1 when do_stage
    const int nb16 = KC / 16;
    const int q = warp & 3, part = (warp - epi_warp0) >> 2;      // TMEM lane quarter, slice phase
    const int Tin = a.in_T > 0 ? a.in_T : T;
    uint32_t a_cnt = 0, it_cnt = 0;
    for (int item = WETTS_BID; item < n_items; item += WETTS_NBLK) {
      int nt, b, t_group0, tiles, nch;
      decode(item, nt, b, t_group0, tiles, nch);
      const long long len = a.lengths ? ldg_i64(a.lengths + b) : (long long)T;
      if (do_stage) {
        const int t_hi = a.in_mask ? (int)(len < Tin ? len : Tin) : Tin;
        const float* in_b = a.in + (long long)b * a.in_bs;
        if (PROF) t_a = clock_now();
        if (p.l2_prefetch) {
          // the activation rows of this CTA's next item and what this item's epilogue reads back, into L2
          const int nxt = item + (int)WETTS_NBLK;
          if (nxt < n_items) {
            int nt_n, b_n, t0_n, tiles_n, nch_n;
            decode(nxt, nt_n, b_n, t0_n, tiles_n, nch_n);
            const long long len_n = a.lengths ? ldg_i64(a.lengths + b_n) : (long long)T;
            const int t_hi_n = a.in_mask ? (int)(len_n < Tin ? len_n : Tin) : Tin;
            const int lo = ep_max(0, t0_n - a.pad_left), hi = ep_min(t_hi_n, t0_n - a.pad_left + tiles_n * MT + (K - 1) * dil);
            l2_prefetch_rows(a.in + (long long)b_n * a.in_bs, a.in_cs, a.Cin, lo, hi, st, n_stagers);
          }
          tc_epilogue_prefetch(a, b, nt * N, ep_min(a.Cout, nt * N + N), t_group0, ep_min(T, t_group0 + tiles * MT), st, n_stagers);
        }
        if (PROF) pc[5] += clock_now() - t_a;
        for (int c = 0; c < nch; ++c) {
          const int c0 = c * KC;
          for (int g = 0; g < tiles; ++g) {
            const uint32_t ab = a_cnt & (NA - 1u), use = a_cnt >> na_log;
            if (PROF) t_a = clock_now();
            uint8_t* Ah = A0 + (size_t)ab * a_bytes;
            const int t_in0 = t_group0 + g * MT - a.pad_left;
            bool waited = (use == 0);
            const int U = (p.debug_skip & 1) ? 0 : nb16 * Rp;
            // two (row, 16-channel) units per round: 32 independent loads in flight per thread
            for (int u = st; u < U; u += 2 * n_stagers) {
              int uu[2] = {u, u + n_stagers};
              float v[2][16];
              int rr[2], qq[2];
              bool has[2];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                has[k] = uu[k] < U;
                qq[k] = uu[k] / Rp;
                rr[k] = uu[k] - qq[k] * Rp;
                const int t = t_in0 + rr[k];
                const bool rok = has[k] && (rr[k] < R) && (t >= 0) && (t < t_hi);
                const int ci0 = c0 + qq[k] * 16;
                const float* src = in_b + (long long)ci0 * a.in_cs + t;
#pragma unroll
                for (int e = 0; e < 16; ++e) v[k][e] = (rok && (ci0 + e) < a.Cin) ? ldg(src + (long long)e * a.in_cs) : 0.f;
              }
              if (!waited) {   // the MMAs that last read this slot must be done before it is overwritten
                if (PROF) t_b = clock_now();
                mbar_wait(bar_a_free + 8 * ab, (use - 1) & 1);
                if (PROF) pc[1] += clock_now() - t_b;
                waited = true;
              }
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                if (has[k]) {
#pragma unroll
                  for (int g8 = 0; g8 < 2; ++g8) {
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                      x[e] = v[k][g8 * 8 + e];
                      if (a.pre_act) x[e] = x[e] > 0.f ? x[e] : x[e] * a.pre_slope;
                    }
                    uint4 hi, lo;
                    f16_split2(x[0], x[1], hi.x, lo.x);
                    f16_split2(x[2], x[3], hi.y, lo.y);
                    f16_split2(x[4], x[5], hi.z, lo.z);
                    f16_split2(x[6], x[7], hi.w, lo.w);
                    const size_t o = ((size_t)(qq[k] * 2 + g8) * Rp + rr[k]) * 16;
                    *reinterpret_cast<uint4*>(Ah + o) = hi;
                    *reinterpret_cast<uint4*>(Ah + a_half + o) = lo;
                  }
                }
              }
            }
            if (!waited) mbar_wait(bar_a_free + 8 * ab, (use - 1) & 1);
            if (PROF) t_b = clock_now();
            fence_async_smem();
            warp_sync();
            if (lane == 0) mbar_arrive(bar_a_full + 8 * ab);
            if (PROF) { const long long t_c = clock_now(); pc[3] += t_c - t_b; pc[2] += t_b - t_a; }   // [2] includes [1]
            a_cnt += 1;
          }
        }
      }
      if (do_epi) {
        const uint32_t slot = it_cnt & s_log, s_use = it_cnt >> s_log;
        if (PROF) t_a = clock_now();
        mbar_wait(bar_acc_full + 8 * slot, s_use & 1);
        if (PROF) { t_b = clock_now(); pc[6] += t_b - t_a; }
        tc_fence_after();
        const int tiles_e = (p.debug_skip & 2) ? 0 : tiles;
        for (int g = 0; g < tiles_e; ++g) {
          for (int mb = 0; mb < MB; ++mb) {
            const int t = t_group0 + g * MT + mb * 128 + q * 32 + lane;
            const float msk = (t < len) ? 1.f : 0.f;
            const uint32_t col0 = slot * slot_cols + (uint32_t)((g * MB + mb) * 2 * N);
            for (int nl = part * 16; nl < N; nl += nparts * 16) {      // 16-column slices, round-robin over the quarter's warps
              float v[16], vs[16];
              tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col0 + (uint32_t)nl, v);
              tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col0 + (uint32_t)(N + nl), vs);
              tmem_ld_wait();
              const int co0 = nt * N + nl;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int co = co0 + i;
                float add = 0.f;
                if (co < a.Cout) {
                  if (a.bias) add = ldg(a.bias + co);
                  if (a.ep.cond) {
                    const float* gp = a.ep.cond + (long long)b * a.ep.cond_bs + a.ep.cond_off;
                    if (a.ep.mode == EPI_GATE) add += (co & 1) ? ldg(gp + a.ep.H + (co >> 1)) : ldg(gp + (co >> 1));
                    else if (a.ep.mode == EPI_PLAIN) add += ldg(gp + co);
                  }
                }
                v[i] = (v[i] + vs[i] * kF16LoInv) + add;
              }
              if (t < T && co0 < a.Cout) tc_epilogue_slice_p(a, b, t, co0, v, msk);
            }
          }
        }
        tc_fence_before();
        warp_sync();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * slot);
        if (PROF) pc[7] += clock_now() - t_b;
        it_cnt += 1;
      }
    }
  }
  if (PROF && p.prof && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 15)) {
    // role rows: 0 MMA issuer, 1 weight producer, 2 first stager warp, 3 last worker warp (drains in both assignments)
    const int role = warp == 15 ? 3 : warp;
    long long* dst = p.prof + ((size_t)WETTS_BID * 4 + role) * (kTc16pProfSlots + 1);
    for (int i = 0; i < kTc16pProfSlots; ++i) dst[i] = pc[i];
    dst[kTc16pProfSlots] = clock_now() - t_start;
  }
  tc_fence_before();
  cta_sync();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace wetts
