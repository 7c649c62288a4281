// Launch side of the fused ResBlock2/MRF stage kernel (fused_rb_kernel.cuh): weight packing into the
// per-item chunk sequence, eligibility checks, persistent-grid launch.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "fused_rb_kernel.cuh"
#include "kernels.cuh"

namespace wetts {
namespace {

// dst[tap][32-channel slice][k-group][hi|lo][n][e] <- folded weight src[co][ci][tap] (3xTF32 split; fused_rb_pack_index)
__global__ void fused_rb_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int K) {
  const long long total = (long long)K * C * C * 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const FusedRbPackIdx ix = fused_rb_pack_index(i, C);
    const float w = src[((long long)ix.co * C + ix.ci) * K + ix.tap];
    const float hi = tc::tf32_rna(w);
    dst[i] = ix.hl ? tc::tf32_rna(w - hi) : hi;
  }
}

// CTA-local shared-window offset of the dynamic shared memory base of a kernel without static shared memory
__global__ void probe_dyn_smem_kernel(uint32_t* out) {
  extern __shared__ __align__(128) uint8_t probe_smem[];
  if (threadIdx.x == 0) *out = tc::smem_u32(probe_smem) & 0xFFFFFFu;
}

}  // namespace

int dyn_smem_offset(uint32_t* off, cudaStream_t s) {
  // per device (the window offset is a property of the device / driver); racing first calls probe twice, harmlessly
  static std::atomic<int> cache[DynSmemAttr::kMaxDev];
  static std::atomic<bool> init{false};
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!init.load()) { for (auto& c : cache) c.store(-1); init.store(true); }
  }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= DynSmemAttr::kMaxDev) return 1;
  int cached = cache[dev].load();
  if (cached < 0) {
    uint32_t* d = nullptr;
    uint32_t h = 0;
    if (cudaMalloc(&d, sizeof(uint32_t)) != cudaSuccess) return 1;
    probe_dyn_smem_kernel<<<1, 32, 1024, s>>>(d);
    if (cudaMemcpyAsync(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(s) != cudaSuccess) return 1;
    cudaFree(d);
    cached = (int)h;
    cache[dev].store(cached);
  }
  *off = (uint32_t)cached;
  return 0;
}

static std::atomic<bool> g_fused_rb{true};
void set_fused_resblock_enabled(bool on) { g_fused_rb = on; }
bool fused_resblock_enabled() { return g_fused_rb; }

bool fused_rb_supported(int C, int nrb, const int* k, const int* d1, const int* d2) {
  if (C != 32 && C != 64) return false;
  if (nrb < 1 || nrb > 3) return false;
  for (int j = 0; j < nrb; ++j) {
    if (k[j] < 1 || (k[j] & 1) == 0) return false;
    const int h1 = d1[j] * (k[j] - 1) / 2, h2 = d2[j] * (k[j] - 1) / 2;
    if (128 + 2 * ((h1 + h2 + 3) & ~3) > kFusedRbPitch) return false;   // rows of the activation tile
    if (2 * h2 > 128) return false;                               // two conv1 blocks must cover x1
  }
  return true;
}

size_t fused_rb_conv_floats(int C, int K) { return (size_t)K * C * C * 2; }

void launch_fused_rb_pack(const float* w_folded, float* dst, int C, int K, cudaStream_t s) {
  const long long total = (long long)fused_rb_conv_floats(C, K);
  const int blocks = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
  fused_rb_pack_kernel<<<blocks, 256, 0, s>>>(w_folded, dst, C, K);
  count_launch();
}

template <int C, int THREADS, int MINB, int NB, bool PROFILE>
static int launch_variant(const FusedRbArgs& a, int grid, size_t smem, cudaStream_t s) {
  auto kern = fused_resblock2_kernel<C, THREADS, MINB, NB, PROFILE>;
  // per launch (not cached): the attribute is per device and costs microseconds
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
  kern<<<grid, THREADS, smem, s>>>(a);
  count_launch();
  return 0;
}
template <bool PROFILE>
static int launch_any(int C, int ring, const FusedRbArgs& a, int grid, size_t smem, cudaStream_t s) {
  if (C == 32) return ring == 6 ? launch_variant<32, 256, 2, 6, PROFILE>(a, grid, smem, s) : launch_variant<32, 256, 2, 4, PROFILE>(a, grid, smem, s);
  if (C == 64) return ring == 6 ? launch_variant<64, 512, 1, 6, PROFILE>(a, grid, smem, s) : launch_variant<64, 512, 1, 4, PROFILE>(a, grid, smem, s);
  return 1;
}

// Debug path (WETTS_FUSED_RB_PROFILE=1): the clock64-instrumented instantiation, phase table to stderr.
static int launch_fused_rb_profiled(int C, int ring, FusedRbArgs a, int grid, size_t smem, long long items, cudaStream_t s) {
  static const char* names[kFusedRbProfPhases] = {"stage", "sync+prefetch", "conv1 issue", "conv1 wait", "epi1", "sync",
                                                  "conv2 issue", "conv2 wait", "epi2", "end sync", "loop", "[conv1 rb0", "conv1 rb1", "conv1 rb2", "full-wait]"};
  const size_t n = (size_t)grid * 2 * kFusedRbProfPhases;
  long long* d = nullptr;
  if (cudaMalloc(&d, n * sizeof(long long)) != cudaSuccess) return 1;
  cudaMemsetAsync(d, 0, n * sizeof(long long), s);
  a.prof = d;
  if (launch_any<true>(C, ring, a, grid, smem, s)) return 1;
  if (cudaStreamSynchronize(s) != cudaSuccess) return 1;
  std::vector<long long> h(n);
  cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  const double per_cta_items = (double)items / grid;
  fprintf(stderr, "[fused_rb profile] C=%d ring=%d B=%d T=%d grid=%d items/CTA=%.1f  (cycles per item, mean over CTAs)\n", C, ring,
          a.B, a.T, grid, per_cta_items);
  for (int who = 0; who < 2; ++who) {
    double tot = 0;
    fprintf(stderr, "  %s:", who ? "thread 32 (producer warp)" : "thread 0 (MMA issuer)   ");
    for (int i = 0; i < kFusedRbProfPhases; ++i) {
      double sum = 0;
      for (int b = 0; b < grid; ++b) sum += (double)h[((size_t)b * 2 + who) * kFusedRbProfPhases + i];
      const double v = sum / grid / per_cta_items;
      if (i < 11) tot += v;
      fprintf(stderr, " %s=%.0f", names[i], v);
    }
    fprintf(stderr, " | total=%.0f\n", tot);
  }
  return 0;
}

int launch_fused_rb(int C, FusedRbArgs a, cudaStream_t s) {
  if ((a.T & 3) != 0 || (((uintptr_t)a.in | (uintptr_t)a.out | (uintptr_t)a.w) & 15) != 0) return 1;   // 16 B loads / bulk copies
  fused_rb_finalize_args(a, C);
  if (dyn_smem_offset(&a.smem_off, s)) return 1;
  const int n_sm = current_device_sm_count();
  if (n_sm <= 0) return 1;
  const char* force = getenv("WETTS_FUSED_RB_RING");
  const int ring = force ? atoi(force) : fused_rb_ring_slots(a.nq);
  if (ring != 4 && !(ring == 6 && a.nq % 6 == 0)) return 1;
  const size_t smem = fused_rb_smem_bytes(C, ring);
  const long long items = (long long)a.B * ((a.T + 127) / 128);
  const int per_sm = (C == 32) ? 2 : 1;
  const int grid = (int)(items < (long long)per_sm * n_sm ? items : (long long)per_sm * n_sm);
  if (getenv("WETTS_FUSED_RB_PROFILE")) return launch_fused_rb_profiled(C, ring, a, grid, smem, items, s);
  return launch_any<false>(C, ring, a, grid, smem, s);
}

int fused_rb_install_fault_word(unsigned int* word) { return tc::install_fault_word_tu(word) == cudaSuccess ? 0 : 1; }

}  // namespace wetts
