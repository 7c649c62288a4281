// Row-block-resident per-layer implicit-GEMM Conv1d on the tensor pipe (sm_100a tcgen05, f16 operand split) for layers
// with several N tiles (C_out > 128): flow in_layer / res_skip, text-encoder and duration-predictor convs, the
// polyphase upsamplers.
//
// conv1d_tc16_kernel (tc16_conv_kernel.cu) treats (N tile, row block) as the work item: the activations of a row block
// are staged -- global fp32 -> f16 split -> shared memory -- once PER N TILE (three times for a flow in_layer, eight times
// for the first upsampler), and an item's phases (stage, MMAs, drain) run one after the other because its accumulators
// fill TMEM.  Measured: those SIMT phases, not the MMAs, bound the kernel (instruction issue).  Here the
// work item is a ROW BLOCK of 128 output rows:
//
//   workers (warps 2..15)  stage the block's activations ONCE, all C_in channels, resident in shared memory
//                          ([chunk][hi | lo'][KC/8][rows][8 halfs], the layout of the MMA A operand);
//   warp 1                 streams the weights of every (N tile, chunk) through a ring (cp.async.bulk + mbarrier);
//   warp 0                 issues the MMAs of N tile 0, 1, 2, .. into the two halves of TMEM in turn (2 x [hi*hi | small]
//                          x N <= 128 columns);
//   workers 4..15          drain N tile i (TMEM -> fused epilogue) while the MMAs of N tile i + 1 run.
//
// Per block: stage -> MMA(0) -> MMA(1) | drain(0) -> MMA(2) | drain(1) -> .. -> drain(last).  Staging instructions drop
// by the number of N tiles, every drain but the last overlaps MMAs; the price is the weight stream of the whole layer per
// 128 rows (what the two-CTA tiling of conv1d_tc16_kernel pays as well).  Same packed weights and plan as that tiling
// (N, KC, chunk count), so a layer can take either route at launch time.
//
// No PTX here (tc_prims.cuh wrappers): the same source runs in the host CTA emulator (tests/emu/tc16r_emu.cpp).
#pragma once
#include "conv_args.h"
#include "tc_epilogue.cuh"
#include "tc_prims.cuh"

namespace wetts {

constexpr int kTc16rThreads = 512;
constexpr int kTc16rNB = 4;              // weight ring slots (maximum)
constexpr int kTc16rWorkerWarp0 = 2, kTc16rWorkers = 14 * 32, kTc16rEpiWarp0 = 4, kTc16rEpiWarps = 12;
constexpr int kTc16rHeader = 192 + 2 * 4 * 1024;   // mbarriers, TMEM slot, two sets of per-item additive terms (<= 1024 output channels)

// shared memory: [header][A resident: n_chunks x (hi | lo')][B ring: NB x weight tile]
inline size_t tc16r_smem_bytes(int K, int dil, int N, int KC, int n_chunks, int nb) {
  const int R = 128 + (K - 1) * dil;
  const int Rp = (R + 7) & ~7;
  return kTc16rHeader + (size_t)n_chunks * (2 * (size_t)(KC / 8) * Rp * 16) + (size_t)nb * ((size_t)K * (KC / 8) * 2 * N * 16);
}

// PROF: clock64 phase counters of one thread per role (WETTS_TC16R_PROFILE=1; table printed by the launcher)
constexpr int kTc16rProfSlots = 8;
template <int MODE, bool PROF = false>
WETTS_GLOBAL void WETTS_LAUNCH_BOUNDS(kTc16rThreads, 1) conv1d_tc16r_kernel(const TcConvArgs p) {
  using namespace tc;
  long long pc[kTc16rProfSlots];
#pragma unroll
  for (int i = 0; i < kTc16rProfSlots; ++i) pc[i] = 0;
  long long t_a = 0, t_b = 0;
  const long long t_start = PROF ? clock_now() : 0;
  WETTS_SMEM_DECL(smem);
  const ConvArgs& a = p.c;
  const int tid = WETTS_TID, lane = tid & 31;
  const int warp = (int)uniform_bits((uint32_t)(tid >> 5), 0, 4);
  const int K = a.K, dil = a.dil, T = a.T;
  const int N = p.N, KC = p.KC, n_chunks = p.n_chunks, n_tiles = p.n_tiles;
  const int R = 128 + (K - 1) * dil;
  const int Rp = p.R_pad;
  const uint32_t a_half = (uint32_t)(KC / 8) * Rp * 16;
  const uint32_t a_bytes = 2 * a_half;
  const uint32_t b_bytes = (uint32_t)K * (KC / 8) * 2 * N * 16;
  const uint32_t NB = (uint32_t)p.n_bbuf;
  const uint32_t slot_cols = (uint32_t)(2 * N);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 184);
  float* addv0 = reinterpret_cast<float*>(smem + 192);         // [2][1024] bias + conditioning of the item (workers are at most one item apart)
  uint8_t* A0 = smem + kTc16rHeader;
  uint8_t* B0 = A0 + (size_t)n_chunks * a_bytes;
  const uint32_t bar_b_full = smem_u32(&bars[0]);     // [4]  bulk copy -> MMA
  const uint32_t bar_b_free = smem_u32(&bars[4]);     // [4]  MMA -> producer
  const uint32_t bar_acc_full = smem_u32(&bars[8]);   // [2]  MMA -> drain warps (and: the block's activations are free)
  const uint32_t bar_acc_empty = smem_u32(&bars[10]); // [2]  drain warps -> MMA
  const uint32_t bar_a_ready = smem_u32(&bars[12]);   //      workers -> MMA: the block is staged
  const uint32_t A_addr = smem_u32(A0), B_addr = smem_u32(B0);

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  if (tid == 0) {
    for (int i = 0; i < kTc16rNB; ++i) {
      mbar_init(bar_b_full + 8 * i, 1);
      mbar_init(bar_b_free + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, kTc16rEpiWarps);
    }
    mbar_init(bar_a_ready, kTc16rWorkers / 32);
    mbar_init_fence();
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = uniform_bits(*tmem_slot, 5, 9);

  const int n_blocks = (T + 127) / 128;
  const int n_items = a.B * n_blocks;
  const int nbulk = (int)((b_bytes + 32767u) / 32768u);

  // item -> (utterance, first row, active): identical arithmetic in every role.  Length-aware mode: a block wholly beyond
  // (len + margin) frames is skipped by every role alike (no barrier ticks).
  auto decode = [&](int item, int& b, int& t0, bool& active) {
    b = item / n_blocks;
    t0 = (item - b * n_blocks) * 128;
    active = true;
    if (a.la_len) active = !((long long)t0 >= (ldg_i64(a.la_len + b) + a.la_margin) * (long long)a.la_rate);
  };

  if (warp == 0) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc_n = idesc_f16_m128(N), idesc_2n = idesc_f16_m128(2 * N);
    const uint32_t a_lo_delta = a_half >> 4;
    uint32_t bb = 0, b_use = 0, acc_cnt = 0, it_cnt = 0;
    for (int item = WETTS_BID; item < n_items; item += WETTS_NBLK) {
      int b, t0;
      bool active;
      decode(item, b, t0, active);
      if (active) {
        if (PROF) t_a = clock_now();
        mbar_wait(bar_a_ready, it_cnt & 1);
        if (PROF) pc[0] += clock_now() - t_a;
        tc_fence_after();
        for (int nt = 0; nt < n_tiles; ++nt) {
          const uint32_t slot = acc_cnt & 1u, s_use = acc_cnt >> 1;
          if (PROF) t_a = clock_now();
          if (s_use > 0) mbar_wait(bar_acc_empty + 8 * slot, (s_use - 1) & 1);   // the slot's previous N tile is drained
          if (PROF) pc[1] += clock_now() - t_a;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + slot * slot_cols;
          for (int c = 0; c < n_chunks; ++c) {
            if (PROF) t_a = clock_now();
            mbar_wait(bar_b_full + 8 * bb, b_use & 1);
            if (PROF) { t_b = clock_now(); pc[2] += t_b - t_a; }
            tc_fence_after();
            const uint64_t adesc0 = make_desc(A_addr + (uint32_t)c * a_bytes, (uint32_t)Rp * 16, 128);
            const uint64_t bdesc0 = make_desc(B_addr + bb * b_bytes, (uint32_t)(2 * N) * 16, 128);
            const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
            for (int tap = 0; tap < K; ++tap) {
              uint32_t al = alo0 + (uint32_t)(tap * dil);
              uint32_t bl = blo0 + (uint32_t)tap * (uint32_t)((KC / 8) * 2 * N);
              for (int kk = 0; kk < KC / 16; ++kk) {
                const uint32_t first = (c == 0 && tap == 0 && kk == 0) ? 0u : 1u;
                tc_mma_f16_split2(d_tmem, d_tmem + (uint32_t)N, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + a_lo_delta),
                                  desc_with_lo(bdesc0, bl), idesc_2n, idesc_n, first);
                al += 2u * (uint32_t)Rp;
                bl += 2u * (uint32_t)(2 * N);
              }
            }
            if (elect_one()) tc_commit(bar_b_free + 8 * bb);
            warp_sync();
            if (PROF) pc[3] += clock_now() - t_b;
            bb += 1;
            if (bb == NB) { bb = 0; b_use += 1; }
          }
          if (elect_one()) tc_commit(bar_acc_full + 8 * slot);
          warp_sync();
          acc_cnt += 1;
        }
        it_cnt += 1;
      }
    }
  } else if (warp == 1) {
    // =============================== weight producer ===============================
    uint32_t bb = 0, use = 0;
    for (int item = WETTS_BID; item < n_items; item += WETTS_NBLK) {
      int b, t0;
      bool active;
      decode(item, b, t0, active);
      if (active) {
        for (int nt = 0; nt < n_tiles; ++nt) {
          for (int c = 0; c < n_chunks; ++c) {
            if (use > 0) mbar_wait(bar_b_free + 8 * bb, (use - 1) & 1);
            warp_sync();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wtc) + ((size_t)nt * n_chunks + c) * b_bytes;
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
      }
    }
  } else {
    // =============================== workers: stage the block, then drain its N tiles ===============================
    const int st = tid - 32 * kTc16rWorkerWarp0;          // 0 .. 447
    const int nb16 = KC / 16;
    const bool do_epi = warp >= kTc16rEpiWarp0;
    const int q = warp & 3, part = (warp - kTc16rEpiWarp0) >> 2;      // TMEM lane quarter, slice phase (0..2)
    const int Tin = a.in_T > 0 ? a.in_T : T;
    const long long in_cs = a.in_cs;
    const float slope = a.pre_act ? a.pre_slope : 1.f;      // lrelu(x) = max(x, slope * x) for slope <= 1; identity at 1
    uint32_t acc_cnt = 0, it_cnt = 0;
    for (int item = WETTS_BID; item < n_items; item += WETTS_NBLK) {
      int b, t0;
      bool active;
      decode(item, b, t0, active);
      if (active) {
        const long long len = a.lengths ? ldg_i64(a.lengths + b) : (long long)T;
        // (the previous block's last N tile was drained -- or, for warps 2 and 3, its accumulators were seen complete --
        // before this point, so its MMAs are done and the resident activations can be overwritten; the additive terms are
        // double buffered because another worker warp may still be draining the previous block)
        if (PROF) t_a = clock_now();
        float* addv = addv0 + (it_cnt & 1u) * 1024;
        for (int n = st; n < n_tiles * N; n += kTc16rWorkers) {
          float x = 0.f;
          if (n < a.Cout) {
            if (a.bias) x = ldg(a.bias + n);
            if (a.ep.cond) {
              const float* gp = a.ep.cond + (long long)b * a.ep.cond_bs + a.ep.cond_off;
              if (MODE == EPI_GATE) x += (n & 1) ? ldg(gp + a.ep.H + (n >> 1)) : ldg(gp + (n >> 1));
              else if (MODE == EPI_PLAIN) x += ldg(gp + n);
            }
          }
          addv[n] = x;
        }
        if (p.l2_prefetch) {
          const int nxt = item + (int)WETTS_NBLK;
          if (nxt < n_items) {
            int b_n, t0_n;
            bool act_n;
            decode(nxt, b_n, t0_n, act_n);
            const long long len_n = a.lengths ? ldg_i64(a.lengths + b_n) : (long long)T;
            const int t_hi_n = a.in_mask ? (int)(len_n < Tin ? len_n : Tin) : Tin;
            const int lo = ep_max(0, t0_n - a.pad_left), hi = ep_min(t_hi_n, t0_n - a.pad_left + R);
            if (act_n) l2_prefetch_rows(a.in + (long long)b_n * a.in_bs, a.in_cs, a.Cin, lo, hi, st, kTc16rWorkers);
          }
          tc_epilogue_prefetch(a, b, 0, a.Cout, t0, ep_min(T, t0 + 128), st, kTc16rWorkers);
        }
        {
          const int t_hi = a.in_mask ? (int)(len < Tin ? len : Tin) : Tin;
          const float* in_b = a.in + (long long)b * a.in_bs;
          const int t_in0 = t0 - a.pad_left;
          const int U = n_chunks * nb16 * Rp;              // (chunk, 16-channel group, row) units of the whole block
          // two units per round: 32 independent loads in flight per thread
          for (int u = st; u < U; u += 2 * kTc16rWorkers) {
            int uu[2] = {u, u + kTc16rWorkers};
            float v[2][16];
            int rr[2], qq[2];
            bool has[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              has[k] = uu[k] < U;
              qq[k] = uu[k] / Rp;                            // 16-channel group over all chunks
              rr[k] = uu[k] - qq[k] * Rp;
              const int t = t_in0 + rr[k];
              const bool rok = has[k] && (rr[k] < R) && (t >= 0) && (t < t_hi);
              const int ci0 = qq[k] * 16;
              const float* sp = in_b + (long long)ci0 * in_cs + t;
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                v[k][e] = (rok && (ci0 + e) < a.Cin) ? ldg(sp) : 0.f;
                sp += in_cs;
              }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (has[k]) {
                const int c = qq[k] / nb16, q16 = qq[k] - c * nb16;
                uint8_t* Ah = A0 + (size_t)c * a_bytes;
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  float x[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) x[e] = fmaxf(v[k][g8 * 8 + e], v[k][g8 * 8 + e] * slope);
                  uint4 hi, lo;
                  f16_split2(x[0], x[1], hi.x, lo.x);
                  f16_split2(x[2], x[3], hi.y, lo.y);
                  f16_split2(x[4], x[5], hi.z, lo.z);
                  f16_split2(x[6], x[7], hi.w, lo.w);
                  const size_t o = ((size_t)(q16 * 2 + g8) * Rp + rr[k]) * 16;
                  *reinterpret_cast<uint4*>(Ah + o) = hi;
                  *reinterpret_cast<uint4*>(Ah + a_half + o) = lo;
                }
              }
            }
          }
          fence_async_smem();
          warp_sync();
          if (lane == 0) mbar_arrive(bar_a_ready);
          if (PROF) pc[4] += clock_now() - t_a;
        }
        for (int nt = 0; nt < n_tiles; ++nt) {
          const uint32_t slot = acc_cnt & 1u, s_use = acc_cnt >> 1;
          if (do_epi) {
            if (PROF) t_a = clock_now();
            mbar_wait(bar_acc_full + 8 * slot, s_use & 1);
            if (PROF) { t_b = clock_now(); pc[5] += t_b - t_a; }
            tc_fence_after();
            const int t = t0 + q * 32 + lane;
            const float msk = (t < len) ? 1.f : 0.f;
            const uint32_t col0 = slot * slot_cols;
            for (int nl = part * 16; nl < N; nl += 3 * 16) {       // 16-column slices, round-robin over the quarter's 3 warps
              float v[16], vs[16];
              tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col0 + (uint32_t)nl, v);
              tmem_ld16_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + col0 + (uint32_t)(N + nl), vs);
              tmem_ld_wait();
              const int co0 = nt * N + nl;
              const float* av = addv + co0;
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = (v[i] + vs[i] * kF16LoInv) + av[i];
              if (t < T && co0 < a.Cout) tc_epilogue_slice_m<MODE>(a, b, t, co0, v, msk);
            }
            tc_fence_before();
            warp_sync();
            if (lane == 0) mbar_arrive(bar_acc_empty + 8 * slot);
            if (PROF) pc[6] += clock_now() - t_b;
          } else {
            // warps 2 and 3 only stage: they may touch the resident activations again once the block's last MMAs are done.
            // They follow every N tile's completion, not only the last one: a parity wait that skips a phase of the
            // barrier returns at once.
            mbar_wait(bar_acc_full + 8 * slot, s_use & 1);
          }
          acc_cnt += 1;
        }
        it_cnt += 1;
      }
    }
  }
  if (PROF && p.prof && lane == 0 && (warp == 0 || warp == 4 || warp == 15)) {
    // rows: 0 MMA issuer, 1 first drain warp (3 slices of 8), 2 last drain warp (2 slices)
    const int role = warp == 0 ? 0 : (warp == 4 ? 1 : 2);
    long long* dst = p.prof + ((size_t)WETTS_BID * 3 + role) * (kTc16rProfSlots + 1);
    for (int i = 0; i < kTc16rProfSlots; ++i) dst[i] = pc[i];
    dst[kTc16rProfSlots] = clock_now() - t_start;
  }
  tc_fence_before();
  cta_sync();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace wetts
