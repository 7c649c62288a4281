// Launch side of the f16-split fused MRF stage kernel (fused_mrf16_kernel.cuh): weight packing into the per-item
// chunk sequence, eligibility checks, persistent-grid launch, optional length-aware tile list.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fused_mrf16_kernel.cuh"
#include "kernels.cuh"

namespace wetts {
namespace {

// dst (halfs) [tap][32-channel slice][4 k-groups][hi | lo'][n][8] <- folded weight src[co][ci][tap]
__global__ void fused_mrf16_pack_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int C, int K) {
  const long long total = (long long)K * C * C * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const FusedMrfPackIdx ix = fused_mrf16_pack_index(i, C);
    const float w = src[((long long)ix.co * C + ix.ci) * K + ix.tap];
    uint32_t hi2, lo2;
    tc::f16_split2(w, 0.f, hi2, lo2);
    dst[i] = (uint16_t)((ix.hl ? lo2 : hi2) & 0xFFFFu);
  }
}

// Work-item list of the length-aware mode: utterance b contributes tiles(b) = ceil(min(T, (len[b] + margin) * rate) / item)
// items of `item` output samples (at least one), in (b, tile) order.  One block: a serial prefix over B (a few thousand at most), then every
// thread fills the segments of its utterances.
__global__ void mrf_item_map_kernel(const long long* __restrict__ lengths, int B, int T, int rate, int margin, int item,
                                    int2_t* __restrict__ item_map, int* __restrict__ n_items, int* __restrict__ prefix) {
  auto tiles = [&](int b) {
    long long n = (lengths[b] + margin) * (long long)rate;
    if (n > T) n = T;
    if (n < 1) n = 1;
    return (int)((n + item - 1) / item);
  };
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) { prefix[b] = acc; acc += tiles(b); }
    prefix[B] = acc;
    *n_items = acc;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int first = prefix[b], n = prefix[b + 1] - first;
    for (int i = 0; i < n; ++i) item_map[first + i] = int2_t{b, i * item};
  }
}

}  // namespace

bool fused_mrf16_supported(int C, int type, int nrb, const int* k, const int (*dil)[kMrfMaxConv], int nconv) {
  if (C != 32 && C != 64 && C != 128) return false;
  if (C == 128 && type != 2) return false;      // ResBlock1 needs two tiles: 2 x 127 KB do not fit
  static const int c128 = getenv("WETTS_MRF16_C128") ? atoi(getenv("WETTS_MRF16_C128")) : 0;   // opt-in until measured
  if (C == 128 && !c128) return false;
  if (nrb < 1 || nrb > kMrfMaxRb) return false;
  if (!((type == 2 && nconv == 2) || (type == 1 && nconv == 6))) return false;
  const int rp = (type == 1) ? 249 : 225;
  for (int j = 0; j < nrb; ++j) {
    if (k[j] < 1 || (k[j] & 1) == 0) return false;
    int H = 0;
    for (int c = 0; c < nconv; ++c) {
      if (dil[j][c] < 1) return false;
      H += dil[j][c] * (k[j] - 1) / 2;
    }
    if (type == 1)
      for (int c = 1; c < nconv; c += 2)
        if (dil[j][c] != 1) return false;                      // decoders.py:118-140: convs2 have dilation 1
    const int Hp = (H + 3) & ~3;
    if (128 + 2 * Hp > rp - 1) return false;                    // rows of the activation tile (QMAX quads)
    const int h0 = dil[j][0] * (k[j] - 1) / 2;
    if (128 + 2 * (H - h0) > 256) return false;                 // two 128-row blocks must cover the first conv's output
  }
  return true;
}

void launch_fused_mrf16_pack(const float* w_folded, void* dst, int C, int K, cudaStream_t s) {
  const long long total = (long long)fused_mrf16_conv_halfs(C, K);
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  fused_mrf16_pack_kernel<<<blocks, 256, 0, s>>>(w_folded, reinterpret_cast<uint16_t*>(dst), C, K);
  count_launch();
}

size_t mrf_item_map_bytes(int B, int T) {
  return sizeof(int2_t) * (size_t)B * ((size_t)(T + 127) / 128) + sizeof(int) * ((size_t)B + 2) + 64;
}
void launch_mrf_item_map(const long long* lengths, int B, int T, int rate, int margin, int item_rows, void* scratch,
                         const int2_t** item_map, const int** n_items_dev, cudaStream_t s) {
  int2_t* map = reinterpret_cast<int2_t*>(scratch);
  int* n_items = reinterpret_cast<int*>(map + (size_t)B * ((size_t)(T + 127) / 128));
  int* prefix = n_items + 1;
  mrf_item_map_kernel<<<1, 256, 0, s>>>(lengths, B, T, rate, margin, item_rows, map, n_items, prefix);
  count_launch();
  *item_map = map;
  *n_items_dev = n_items;
}

// Output samples per work item (fused_mrf16_kernel's ITEM).  Larger items amortise the halo: per 256 output samples a
// ResBlock2 conv pair costs 6 M blocks with 128-sample items, 5 with 256, 4.67 with 384, and the staged rows drop by 13 /
// 18 %; the price is one more accumulator block in TMEM per 128 samples (C = 32: 128 / 256 / 256 columns -> three / two / two
// CTAs per SM; C = 64: 256 / 512 columns -> two / one CTA per SM).
// Measured on one B200 (same box, two repeats each, profiles/r02w_*, r02x_*): C = 32 ResBlock2 with 384-sample items
// -2.0 ms of generator time per step (53.3 -> 51.3 ms), 256-sample items -0.8 ms; C = 64 with 256-sample items (one CTA per
// SM) +0.8 ms; ResBlock1 C = 32 with 256-sample items (one CTA per SM: 124 KB) +0.3 ms.  Hence: 384 for the C = 32
// ResBlock2 stage when the launch has at least four items per CTA slot (a small launch -- B = 1 -- is latency bound and
// keeps the short items: more CTAs busy, shorter critical path), 128 everywhere else.
// WETTS_MRF16_ITEM_C32 / _C64 / _RB1 = 128 | 256 (| 384 for C32) force a size per kernel family (experiments).
// process-wide option "mrf_item_rows" (wetts_set_option): 0 = the policy above, 128 / 256 / 384 = that size for the C = 32
// ResBlock2 stage (tests compare the sizes bit for bit); "mrf_item_rows_last" reads back what the last such launch used
static std::atomic<int> g_item_opt{0}, g_item_last{0};
int set_mrf16_item_rows(int rows) {
  if (rows != 0 && rows != 128 && rows != 256 && rows != 384) return 1;
  g_item_opt.store(rows);
  return 0;
}
int mrf16_item_rows_option() { return g_item_opt.load(); }
int mrf16_last_item_rows() { return g_item_last.load(); }

int fused_mrf16_item_rows(int C, int type, int B, int T) {
  static const int c32_env = getenv("WETTS_MRF16_ITEM_C32") ? atoi(getenv("WETTS_MRF16_ITEM_C32")) : 0;
  const int c32 = g_item_opt.load() ? g_item_opt.load() : c32_env;
  static const int c64 = getenv("WETTS_MRF16_ITEM_C64") ? atoi(getenv("WETTS_MRF16_ITEM_C64")) : 128;
  static const int rb1 = getenv("WETTS_MRF16_ITEM_RB1") ? atoi(getenv("WETTS_MRF16_ITEM_RB1")) : 128;
  if (type == 2 && C == 32) {
    if (c32 == 128 || c32 == 256 || c32 == 384) return c32;
    const int n_sm = current_device_sm_count();
    const long long n384 = (long long)B * ((T + 383) / 384);
    return (n_sm > 0 && n384 >= 8LL * n_sm) ? 384 : 128;       // two CTAs per SM: >= 4 items per CTA slot
  }
  if (type == 2 && C == 64) return c64 == 256 ? 256 : 128;
  if (type == 1 && C == 32) return rb1 == 256 ? 256 : 128;
  return 128;
}
static int tile_pitch(int type, int item) {
  if (item == 384) return 481;                     // ResBlock2, C = 32 only: four accumulator blocks = 256 TMEM columns
  return item == 256 ? (type == 1 ? 377 : 353) : (type == 1 ? 249 : 225);
}

template <int C, int THREADS, int MINB, int NB, int RP, bool TWO, bool PROFILE, int ITEM = 128>
static int launch_variant(const FusedMrfArgs& a, int grid, size_t smem, cudaStream_t s) {
  auto kern = fused_mrf16_kernel<C, THREADS, MINB, NB, RP, TWO, PROFILE, ITEM>;
  static DynSmemAttr attr;
  if (attr.ensure((const void*)kern, smem) != cudaSuccess) return 1;
  kern<<<grid, THREADS, smem, s>>>(a);
  count_launch();
  return 0;
}

// CTAs per SM: the accumulators are reused by every conv (4N TMEM columns) and the f16 tiles are half the size of the
// tf32 ones, so the ResBlock2 kernels fit 3 (C = 32) / 2... CTAs; co-resident CTAs overlap one CTA's SIMT phases with
// another's MMAs.  WETTS_MRF16_CTAS overrides (experiments).
// C = 64, ResBlock2: two 256-thread CTAs per SM instead of one 512-thread CTA: -0.9 ms per step in a same-box A/B (two
// repeats each: 73.35 / 73.57 -> 72.46 / 72.64 ms, profiles/r02t_fused_variants_same_box_ab.txt); WETTS_MRF16_C64_CTAS=1 restores it
static const int g_c64_ctas = getenv("WETTS_MRF16_C64_CTAS") ? atoi(getenv("WETTS_MRF16_C64_CTAS")) : 2;
static int ctas_per_sm(int C, int type) {
  static const int forced = getenv("WETTS_MRF16_CTAS") ? atoi(getenv("WETTS_MRF16_CTAS")) : 0;
  if (forced > 0) return forced;
  if (C == 32) return type == 2 ? 3 : 2;
  if (C == 128) return 1;
  return (type == 2) ? g_c64_ctas : 1;
}

template <bool PROFILE>
static int launch_any(int C, int type, int ring, int per_sm, int item, const FusedMrfArgs& a, int grid, size_t smem, cudaStream_t s) {
#define V(CC, TH, MB, NBB, RPP, TW) launch_variant<CC, TH, MB, NBB, RPP, TW, PROFILE>(a, grid, smem, s)
#define V256(CC, TH, MB, NBB, RPP, TW) launch_variant<CC, TH, MB, NBB, RPP, TW, PROFILE, 256>(a, grid, smem, s)
  if (item == 384) {
    if (type == 2 && C == 32)
      return ring == 6 ? launch_variant<32, 256, 2, 6, 481, false, PROFILE, 384>(a, grid, smem, s)
                       : launch_variant<32, 256, 2, 4, 481, false, PROFILE, 384>(a, grid, smem, s);
    return 1;
  }
  if (item == 256) {
    if (type == 2 && C == 32) return ring == 6 ? V256(32, 256, 2, 6, 353, false) : V256(32, 256, 2, 4, 353, false);
    if (type == 2 && C == 64) return ring == 6 ? V256(64, 512, 1, 6, 353, false) : V256(64, 512, 1, 4, 353, false);
    if (type == 1 && C == 32) return ring == 6 ? V256(32, 256, 1, 6, 377, true) : V256(32, 256, 1, 4, 377, true);
    return 1;
  }
  if (type == 2 && C == 32) {
    if (per_sm >= 3) return ring == 6 ? V(32, 256, 3, 6, 225, false) : V(32, 256, 3, 4, 225, false);
    return ring == 6 ? V(32, 256, 2, 6, 225, false) : V(32, 256, 2, 4, 225, false);
  }
  if (type == 2 && C == 128) return ring == 6 ? V(128, 512, 1, 6, 225, false) : V(128, 512, 1, 4, 225, false);
  if (type == 2 && C == 64 && per_sm >= 2) return ring == 6 ? V(64, 256, 2, 6, 225, false) : V(64, 256, 2, 4, 225, false);
  if (type == 2 && C == 64) return ring == 6 ? V(64, 512, 1, 6, 225, false) : V(64, 512, 1, 4, 225, false);
  if (type == 1 && C == 32) return ring == 6 ? V(32, 256, 2, 6, 249, true) : V(32, 256, 2, 4, 249, true);
  if (type == 1 && C == 64) return ring == 6 ? V(64, 512, 1, 6, 249, true) : V(64, 512, 1, 4, 249, true);
#undef V
#undef V256
  return 1;
}

static int launch_profiled(int C, int type, int ring, int per_sm, int item, FusedMrfArgs a, int grid, size_t smem, long long items,
                           cudaStream_t s) {
  static const char* names[kMrfProfPhases] = {"stage", "sync", "conv issue|prefetch", "acc wait", "epilogue", "sync", "out",
                                              "[full-wait]", "loop", "-", "-", "-"};
  const size_t n = (size_t)grid * 2 * kMrfProfPhases;
  long long* d = nullptr;
  if (cudaMalloc(&d, n * sizeof(long long)) != cudaSuccess) return 1;
  cudaMemsetAsync(d, 0, n * sizeof(long long), s);
  a.prof = d;
  if (launch_any<true>(C, type, ring, per_sm, item, a, grid, smem, s)) return 1;
  if (cudaStreamSynchronize(s) != cudaSuccess) return 1;
  std::vector<long long> h(n);
  cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  const double per_cta_items = (double)items / grid;
  fprintf(stderr, "[fused_mrf16 profile] type=%d C=%d ring=%d ctas/sm=%d item=%d B=%d T=%d grid=%d items/CTA=%.1f  (cycles per item, mean over CTAs)\n",
          type, C, ring, per_sm, item, a.B, a.T, grid, per_cta_items);
  for (int who = 0; who < 2; ++who) {
    double tot = 0;
    fprintf(stderr, "  %s:", who ? "thread 32 (producer warp)" : "thread 0 (MMA issuer)   ");
    for (int i = 0; i < 9; ++i) {
      double sum = 0;
      for (int b = 0; b < grid; ++b) sum += (double)h[((size_t)b * 2 + who) * kMrfProfPhases + i];
      const double v = sum / grid / per_cta_items;
      if (i != 7) tot += v;
      fprintf(stderr, " %s=%.0f", names[i], v);
    }
    fprintf(stderr, " | total=%.0f\n", tot);
  }
  return 0;
}

int launch_fused_mrf16(int C, FusedMrfArgs a, cudaStream_t s) {
  if ((a.T & 3) != 0 || (((uintptr_t)a.in | (uintptr_t)a.out | (uintptr_t)a.w) & 15) != 0) return 1;   // 16 B loads / bulk copies
  fused_mrf16_finalize_args(a, C);
  if (dyn_smem_offset(&a.smem_off, s)) return 1;
  const int n_sm = current_device_sm_count();
  if (n_sm <= 0) return 1;
  const char* force = getenv("WETTS_FUSED_RB_RING");
  int ring = force ? atoi(force) : fused_mrf16_ring_slots(a.nq);
  const int item = fused_mrf16_item_rows(C, a.type, a.B, a.T);
  if (C == 32 && a.type == 2) g_item_last.store(item);
  const int rp = tile_pitch(a.type, item);
  if (ring == 6 && fused_mrf16_smem_bytes(C, 6, rp, a.type == 1 ? 2 : 1) > 227 * 1024) ring = 4;
  if (ring != 4 && !(ring == 6 && a.nq % 6 == 0)) return 1;
  const size_t smem = fused_mrf16_smem_bytes(C, ring, rp, a.type == 1 ? 2 : 1);
  if (smem > 227 * 1024) return 1;
  const long long items = (long long)a.B * ((a.T + item - 1) / item);   // upper bound in the length-aware mode
  // 256-sample items: three accumulator blocks -> 256 (C = 32) / 512 (C = 64) TMEM columns per CTA
  const int per_sm = item >= 256 ? ((C == 32 && a.type == 2) ? 2 : 1) : ctas_per_sm(C, a.type);
  static const int stagger = getenv("WETTS_MRF16_STAGGER") ? atoi(getenv("WETTS_MRF16_STAGGER")) : 0;
  a.stagger = per_sm > 1 ? stagger : 0;
  a.n_sm = n_sm;
  const int grid = (int)(items < (long long)per_sm * n_sm ? items : (long long)per_sm * n_sm);
  if (getenv("WETTS_FUSED_RB_PROFILE")) return launch_profiled(C, a.type, ring, per_sm, item, a, grid, smem, items, s);
  return launch_any<false>(C, a.type, ring, per_sm, item, a, grid, smem, s);
}

int fused_mrf16_install_fault_word(unsigned int* word) { return tc::install_fault_word_tu(word) == cudaSuccess ? 0 : 1; }

}  // namespace wetts
