// Launch side of the fused ResBlock2/MRF stage kernel (fused_rb_kernel.cuh): weight packing into the
// per-item chunk sequence, eligibility checks, persistent-grid launch.
#include <cstdio>

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

}  // namespace

static bool g_fused_rb = true;
void set_fused_resblock_enabled(bool on) { g_fused_rb = on; }
bool fused_resblock_enabled() { return g_fused_rb; }

bool fused_rb_supported(int C, int nrb, const int* k, const int* d1, const int* d2) {
  if (C != 32 && C != 64) return false;
  if (nrb < 1 || nrb > 3) return false;
  for (int j = 0; j < nrb; ++j) {
    if (k[j] < 1 || (k[j] & 1) == 0) return false;
    const int h1 = d1[j] * (k[j] - 1) / 2, h2 = d2[j] * (k[j] - 1) / 2;
    if (128 + 2 * (h1 + h2) > 32 * kFusedRbUnits) return false;   // staging units per thread
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

int launch_fused_rb(int C, FusedRbArgs a, cudaStream_t s) {
  fused_rb_finalize_args(a, C);
  const size_t smem = fused_rb_smem_bytes(C, a.Rp);
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long items = (long long)a.B * ((a.T + 127) / 128);
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
