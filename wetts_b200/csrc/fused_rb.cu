// Launch side of the fused ResBlock2/MRF stage kernel (fused_rb_kernel.cuh): weight packing into the
// per-item chunk sequence, eligibility checks, persistent-grid launch.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fused_rb_kernel.cuh"
#include "kernels.cuh"

namespace wetts {
namespace {

// dst[tap][kh][hi|lo][kg][n][e] <- folded weight src[co][ci][tap] (3xTF32 split)
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
  static int cached = -1;
  if (cached < 0) {
    uint32_t* d = nullptr;
    uint32_t h = 0;
    if (cudaMalloc(&d, sizeof(uint32_t)) != cudaSuccess) return 1;
    probe_dyn_smem_kernel<<<1, 32, 1024, s>>>(d);
    if (cudaMemcpyAsync(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess) return 1;
    if (cudaStreamSynchronize(s) != cudaSuccess) return 1;
    cudaFree(d);
    cached = (int)h;
  }
  *off = (uint32_t)cached;
  return 0;
}

static bool g_fused_rb = true;
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

// Debug path (WETTS_FUSED_RB_PROFILE=1): the clock64-instrumented instantiation, phase table to stderr.
static int launch_fused_rb_profiled(int C, FusedRbArgs a, size_t smem, int n_sm, long long items, cudaStream_t s) {
  static const char* names[kFusedRbProfPhases] = {"stage", "sync+prefetch", "conv1 issue", "conv1 wait", "epi1", "sync",
                                                  "conv2 issue", "conv2 wait", "epi2", "end sync", "loop", "[conv1 rb0", "conv1 rb1", "conv1 rb2", "full-wait]"};
  const int grid = (int)(C == 32 ? (items < 2 * n_sm ? items : 2 * n_sm) : (items < n_sm ? items : n_sm));
  const size_t n = (size_t)grid * 2 * kFusedRbProfPhases;
  long long* d = nullptr;
  if (cudaMalloc(&d, n * sizeof(long long)) != cudaSuccess) return 1;
  cudaMemsetAsync(d, 0, n * sizeof(long long), s);
  a.prof = d;
  if (C == 32) {
    cudaFuncSetAttribute(fused_resblock2_kernel<32, 256, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fused_resblock2_kernel<32, 256, 2, true><<<grid, 256, smem, s>>>(a);
  } else {
    cudaFuncSetAttribute(fused_resblock2_kernel<64, 512, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fused_resblock2_kernel<64, 512, 1, true><<<grid, 512, smem, s>>>(a);
  }
  count_launch();
  if (cudaStreamSynchronize(s) != cudaSuccess) return 1;
  std::vector<long long> h(n);
  cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  const double per_cta_items = (double)items / grid;
  fprintf(stderr, "[fused_rb profile] C=%d B=%d T=%d grid=%d items/CTA=%.1f  (cycles per item, mean over CTAs)\n", C, a.B, a.T, grid,
          per_cta_items);
  for (int who = 0; who < 2; ++who) {
    double tot = 0;
    fprintf(stderr, "  %s:", who ? "thread 32 (epilogue warp)" : "thread 0 (MMA issuer)   ");
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
  fused_rb_finalize_args(a, C);
  if (dyn_smem_offset(&a.smem_off, s)) return 1;
  const size_t smem = fused_rb_smem_bytes(C, a.Rp);
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long items = (long long)a.B * ((a.T + 127) / 128);
  if (getenv("WETTS_FUSED_RB_PROFILE")) return launch_fused_rb_profiled(C, a, smem, n_sm, items, s);
  static size_t configured[2] = {0, 0};
  if (C == 32) {
    if (smem > configured[0]) {
      if (cudaFuncSetAttribute(fused_resblock2_kernel<32, 256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
      configured[0] = smem;
    }
    const int grid = (int)(items < 2 * n_sm ? items : 2 * n_sm);
    fused_resblock2_kernel<32, 256, 2><<<grid, 256, smem, s>>>(a);
  } else if (C == 64) {
    if (smem > configured[1]) {
      if (cudaFuncSetAttribute(fused_resblock2_kernel<64, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
      configured[1] = smem;
    }
    const int grid = (int)(items < n_sm ? items : n_sm);
    fused_resblock2_kernel<64, 512, 1><<<grid, 512, smem, s>>>(a);
  } else {
    return 1;
  }
  count_launch();
  return 0;
}

}  // namespace wetts
