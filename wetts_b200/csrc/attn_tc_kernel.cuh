// Windowed relative-position multi-head attention of the text encoder on the tensor pipe (sm_100a tcgen05, f16 split).
//
//   scores = (q / sqrt(dk)) k^T  + band bias (q / sqrt(dk)) . E_k[j - i + w],  |j - i| <= w      attentions.py:244-259
//   scores[mask == 0] = -1e4 ; p = softmax(scores)                                               attentions.py:262-263
//   out = p v + sum_{|j-i|<=w} p[i, j] E_v[j - i + w]                                             attentions.py:272-279
//
// One CTA = (utterance b, head h) with up to 128 queries and up to 128 keys (the benchmark's Tx = 128; longer texts
// take the fp32 SIMT kernel).  Both contractions run as tcgen05.mma kind::f16 with the 2^11-scaled f16 operand split
// of tc_prims.cuh (22 significand bits, two MMAs per k-step, small terms in their own accumulator columns):
//   S[128 x Tk]  : A = q (rows = queries, K = dk channels), B = k (rows = keys; hi rows then lo' rows), 6 k-steps
//   O[128 x dk]  : A = p (rows = queries, K = keys), B = v^T (rows = channels; hi then lo'), Tk/16 k-steps
// q, k, v are staged from the fused qkv conv output [B][3C][T] with the fp32 -> f16 split done on the way; the scores
// never leave the SM: the softmax reads them from TMEM (one query row per thread, three passes: max, sum, write),
// adds the banded key bias, applies the masks and writes p straight into the A-operand layout of the second MMA.
// The banded value term is 9 FMAs per output channel in the O epilogue.
//
// No PTX here: hardware primitives are the tc_prims.cuh wrappers, so the same source runs in the host CTA emulator
// (tests/emu/attn_tc_emu.cpp).
#pragma once
#include <math.h>

#include "tc_prims.cuh"

namespace wetts {

struct AttnTcArgs {
  const float* qkv = nullptr;      // [B][3C][T]: q rows 0..C-1, k rows C..2C-1, v rows 2C..3C-1
  const float* emb_k = nullptr;    // [2w+1][dk]
  const float* emb_v = nullptr;    // [2w+1][dk]
  const long long* lengths = nullptr;
  float* out = nullptr;            // [B][C][T]
  int B = 0, C = 0, T = 0, n_heads = 0, window = 4;
  uint32_t smem_off = 0;           // CTA-local shared-window offset of the dynamic shared memory base
};

constexpr int kAttnTcMaxT = 128;    // queries and keys per CTA
constexpr int kAttnTcDk = 96;       // head dimension this instantiation is built for (192 channels / 2 heads)
constexpr int kAttnTcThreads = 128;

// shared memory: [bars 64 B][Q hi|lo][K hi+lo' rows][P hi|lo][V^T hi+lo' rows]
constexpr uint32_t kAttnQHalf = (kAttnTcDk / 8) * 128 * 16;             // 24 KB
constexpr uint32_t kAttnKBytes = (kAttnTcDk / 8) * (2 * 128) * 16;      // 48 KB   [dk/8][hi keys | lo' keys][8]
constexpr uint32_t kAttnPHalf = (kAttnTcMaxT / 8) * 128 * 16;           // 32 KB
constexpr uint32_t kAttnVBytes = (kAttnTcMaxT / 8) * (2 * kAttnTcDk) * 16;   // 48 KB [Tk/8][hi ch | lo' ch][8]
constexpr uint32_t kAttnTcSmem = 64 + 2 * kAttnQHalf + kAttnKBytes + 2 * kAttnPHalf + kAttnVBytes;
// SHARED = two CTAs per SM: shared memory is used twice -- [q | k] (96 KB) for S = q k^T, then, q and k being dead once the S
// MMAs have completed, [p (64 KB) | v (48 KB)] for O = p v -- and O reuses the TMEM columns of S: 112 KB and 256 columns per
// CTA.  The kernel is a chain of latency-bound phases (staging, softmax from TMEM, epilogue) on four warps; a second
// resident CTA is what fills the SM.
constexpr uint32_t kAttnTcSmemShared = 64 + 2 * kAttnPHalf + kAttnVBytes;
static_assert(2 * kAttnQHalf + kAttnKBytes <= 2 * kAttnPHalf + kAttnVBytes, "q | k must fit in the p | v region");

template <bool SHARED>
WETTS_GLOBAL void WETTS_LAUNCH_BOUNDS(kAttnTcThreads, (SHARED ? 2 : 1)) rel_attention_tc_kernel(const AttnTcArgs p) {
  constexpr uint32_t TMEM_COLS = SHARED ? 256u : 512u;
  using namespace tc;
  constexpr int DK = kAttnTcDk, TK = kAttnTcMaxT, W = 4, NREL = 2 * W + 1;
  WETTS_SMEM_DECL(smem);
  const int tid = WETTS_TID, lane = tid & 31;
  const int warp = (int)uniform_bits((uint32_t)(tid >> 5), 0, 2);
  const int bh = WETTS_BID, b = bh / p.n_heads, h = bh - b * p.n_heads;
  const int T = p.T, C = p.C;
  const long long len = p.lengths[b];
  const float scale = 1.0f / sqrtf((float)DK);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32);
  uint8_t* Qs = smem + 64;
  uint8_t* Ks = Qs + 2 * kAttnQHalf;
  uint8_t* Ps = SHARED ? smem + 64 : Ks + kAttnKBytes;   // SHARED: overlays q | k (written after the S MMAs have completed)
  uint8_t* Vs = Ps + 2 * kAttnPHalf;                     // SHARED: overlays the tail of k (staged after the S MMAs)
  const uint32_t bar_s = smem_u32(&bars[0]), bar_o = smem_u32(&bars[1]);
  if ((smem_u32(smem) & 0xFFFFFFu) != p.smem_off) trap_now();     // same value in every thread: a uniform branch
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init_fence();
  }
  const float* qb = p.qkv + ((long long)b * 3 * C + h * DK) * T;
  const float* kb = qb + (long long)C * T;
  const float* vb = kb + (long long)C * T;

  // ---- stage q (scaled) and k: thread = one time row, 8-channel groups -> one 16 B store per half
  {
    const int t = tid;
    const bool ok = t < T;
    for (int cg = 0; cg < DK / 8; ++cg) {
      float q8[8], k8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        q8[e] = ok ? ldg(qb + (long long)(cg * 8 + e) * T + t) * scale : 0.f;
        k8[e] = ok ? ldg(kb + (long long)(cg * 8 + e) * T + t) : 0.f;
      }
      uint4 hi, lo;
      f16_split2(q8[0], q8[1], hi.x, lo.x); f16_split2(q8[2], q8[3], hi.y, lo.y);
      f16_split2(q8[4], q8[5], hi.z, lo.z); f16_split2(q8[6], q8[7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(Qs + ((size_t)cg * 128 + t) * 16) = hi;
      *reinterpret_cast<uint4*>(Qs + kAttnQHalf + ((size_t)cg * 128 + t) * 16) = lo;
      f16_split2(k8[0], k8[1], hi.x, lo.x); f16_split2(k8[2], k8[3], hi.y, lo.y);
      f16_split2(k8[4], k8[5], hi.z, lo.z); f16_split2(k8[6], k8[7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(Ks + ((size_t)cg * 256 + t) * 16) = hi;           // rows 0..127: hi
      *reinterpret_cast<uint4*>(Ks + ((size_t)cg * 256 + 128 + t) * 16) = lo;     // rows 128..255: lo'
    }
  }
  // ---- stage v^T: element (channel d, key j) at (j/8)*(2*DK*16) + d*16 + (j%8)*2 ; thread = (channel, 8-key group)
  auto stage_v = [&]() {
    for (int u = tid; u < DK * (TK / 8); u += kAttnTcThreads) {
      const int d = u % DK, jg = u / DK;
      float v8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = jg * 8 + e;
        v8[e] = (j < T) ? ldg(vb + (long long)d * T + j) : 0.f;
      }
      uint4 hi, lo;
      f16_split2(v8[0], v8[1], hi.x, lo.x); f16_split2(v8[2], v8[3], hi.y, lo.y);
      f16_split2(v8[4], v8[5], hi.z, lo.z); f16_split2(v8[6], v8[7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(Vs + ((size_t)jg * (2 * DK) + d) * 16) = hi;
      *reinterpret_cast<uint4*>(Vs + ((size_t)jg * (2 * DK) + DK + d) * 16) = lo;
    }
  };
  if (!SHARED) stage_v();
  fence_async_smem();
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = uniform_bits(*tmem_slot, 5, 9);

  // ---- S = q k^T : accumulator columns [0, 128) main, [128, 256) small terms
  constexpr uint32_t Q_OFF = 64u, K_OFF = Q_OFF + 2u * kAttnQHalf, P_OFF = SHARED ? 64u : K_OFF + kAttnKBytes, V_OFF = P_OFF + 2u * kAttnPHalf;
  if (warp == 0) {
    const uint64_t qd = make_desc(p.smem_off + Q_OFF, 128u * 16u, 128u);
    const uint64_t kd = make_desc(p.smem_off + K_OFF, 256u * 16u, 128u);
    const uint32_t idesc_n = idesc_f16_m128(TK), idesc_2n = idesc_f16_m128(2 * TK);
#pragma unroll
    for (int kk = 0; kk < DK / 16; ++kk) {
      const uint32_t al = (uint32_t)qd + (uint32_t)(kk * 2 * 128), bl = (uint32_t)kd + (uint32_t)(kk * 2 * 256);
      tc_mma_f16_split2(tmem_base, tmem_base + (uint32_t)TK, desc_with_lo(qd, al), desc_with_lo(qd, al + (kAttnQHalf >> 4)),
                        desc_with_lo(kd, bl), idesc_2n, idesc_n, kk == 0 ? 0u : 1u);
    }
    if (elect_one()) tc_commit(bar_s);
    warp_sync();
  }
  // while the tensor pipe works: this thread's banded key bias  q_i . E_k[r]  (q re-read from global: L1/L2 hits)
  const int i = tid;                      // query row of this thread = TMEM lane
  const bool row_ok = i < T;
  float bias_r[NREL];
#pragma unroll
  for (int r = 0; r < NREL; ++r) bias_r[r] = 0.f;
  if (row_ok) {
    for (int d = 0; d < DK; ++d) {
      const float qv = ldg(qb + (long long)d * T + i) * scale;
#pragma unroll
      for (int r = 0; r < NREL; ++r) bias_r[r] = fmaf(qv, ldg(p.emb_k + r * DK + d), bias_r[r]);
    }
  }
  mbar_wait(bar_s, 0);
  tc_fence_after();
  if (SHARED) stage_v();   // every S MMA has completed: q and k are dead, their shared memory takes v and p

  // ---- masked softmax over the keys of row i, scores read from TMEM in 16-column slices (three passes)
  const uint32_t lane_sel = (uint32_t)(32 * warp) << 16;
  const bool qvalid = (long long)i < len;
  auto score16 = [&](int j0, float* s) {   // scores of keys j0 .. j0+15 incl. band bias and masks (-inf beyond T)
    float v[16], vs[16];
    tmem_ld16_nowait(tmem_base + lane_sel + (uint32_t)j0, v);
    tmem_ld16_nowait(tmem_base + lane_sel + (uint32_t)(TK + j0), vs);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int j = j0 + e;
      float x = v[e] + vs[e] * kF16LoInv;
      const int r = j - i + W;
      if (r >= 0 && r < NREL) {
        // select without dynamic register indexing
        float br = 0.f;
#pragma unroll
        for (int rr = 0; rr < NREL; ++rr) br = (rr == r) ? bias_r[rr] : br;
        x += br;
      }
      if (!qvalid || (long long)j >= len) x = -1e4f;
      s[e] = (j < T) ? x : -INFINITY;
    }
  };
  float mx = -INFINITY;
  for (int j0 = 0; j0 < TK; j0 += 16) {
    float s[16];
    score16(j0, s);
#pragma unroll
    for (int e = 0; e < 16; ++e) mx = fmaxf(mx, s[e]);
  }
  float sum = 0.f;
  for (int j0 = 0; j0 < TK; j0 += 16) {
    float s[16];
    score16(j0, s);
#pragma unroll
    for (int e = 0; e < 16; ++e) sum += expf(s[e] - mx);
  }
  const float inv = 1.0f / sum;
  float p_band[NREL];
#pragma unroll
  for (int r = 0; r < NREL; ++r) p_band[r] = 0.f;
  for (int j0 = 0; j0 < TK; j0 += 16) {
    float s[16];
    score16(j0, s);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      s[e] = row_ok ? expf(s[e] - mx) * inv : 0.f;
      const int r = j0 + e - i + W;
#pragma unroll
      for (int rr = 0; rr < NREL; ++rr) p_band[rr] = (rr == r) ? s[e] : p_band[rr];
    }
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      uint4 hi, lo;
      f16_split2(s[8 * g8 + 0], s[8 * g8 + 1], hi.x, lo.x); f16_split2(s[8 * g8 + 2], s[8 * g8 + 3], hi.y, lo.y);
      f16_split2(s[8 * g8 + 4], s[8 * g8 + 5], hi.z, lo.z); f16_split2(s[8 * g8 + 6], s[8 * g8 + 7], hi.w, lo.w);
      const size_t o = ((size_t)(j0 / 8 + g8) * 128 + i) * 16;
      *reinterpret_cast<uint4*>(Ps + o) = hi;
      *reinterpret_cast<uint4*>(Ps + kAttnPHalf + o) = lo;
    }
  }
  fence_async_smem();
  tc_fence_before();
  cta_sync();
  tc_fence_after();

  // ---- O = p v : accumulator columns [O_COL, O_COL+DK) main, [O_COL+DK, O_COL+2DK) small terms (SHARED: the columns of S --
  // every thread has read its scores for the last time before the CTA barrier above)
  constexpr uint32_t O_COL = SHARED ? 0u : 256u;
  if (warp == 0) {
    const uint64_t pd = make_desc(p.smem_off + P_OFF, 128u * 16u, 128u);
    const uint64_t vd = make_desc(p.smem_off + V_OFF, (uint32_t)(2 * DK) * 16u, 128u);
    const uint32_t idesc_n = idesc_f16_m128(DK), idesc_2n = idesc_f16_m128(2 * DK);
#pragma unroll
    for (int kk = 0; kk < TK / 16; ++kk) {
      const uint32_t al = (uint32_t)pd + (uint32_t)(kk * 2 * 128), bl = (uint32_t)vd + (uint32_t)(kk * 2 * 2 * DK);
      tc_mma_f16_split2(tmem_base + O_COL, tmem_base + O_COL + (uint32_t)DK, desc_with_lo(pd, al),
                        desc_with_lo(pd, al + (kAttnPHalf >> 4)), desc_with_lo(vd, bl), idesc_2n, idesc_n, kk == 0 ? 0u : 1u);
    }
    if (elect_one()) tc_commit(bar_o);
    warp_sync();
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  // ---- epilogue: + banded value term, store [B][C][T] (lanes across time: coalesced per channel)
  float* ob = p.out + ((long long)b * C + h * DK) * T + i;
  for (int d0 = 0; d0 < DK; d0 += 16) {
    float v[16], vs[16];
    tmem_ld16_nowait(tmem_base + lane_sel + O_COL + (uint32_t)d0, v);
    tmem_ld16_nowait(tmem_base + lane_sel + O_COL + (uint32_t)(DK + d0), vs);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float o = v[e] + vs[e] * kF16LoInv;
#pragma unroll
        for (int r = 0; r < NREL; ++r) o = fmaf(p_band[r], ldg(p.emb_v + r * DK + d0 + e), o);
        ob[(long long)(d0 + e) * T] = o;
      }
    }
  }
  tc_fence_before();
  cta_sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace wetts
