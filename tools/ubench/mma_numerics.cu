// Hardware probe (round 2): (1) numerical error of split-operand tcgen05 GEMMs against an fp64 reference as a
// function of the number of sequential accumulations, for the operand splits the kernels use or may use:
//     tf32 x3 (one accumulator), tf32 split2 ([hi*hi | small terms]), f16 split2 with the lo parts scaled by 2^11,
//     and the same with the k-steps dealt round-robin over 2 / 4 accumulator sets;
// (2) sustained cycles per MMA pair of the fused kernels' issue pattern for kind::tf32 (K = 8) and kind::f16 (K = 16).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I wetts_b200/csrc tools/ubench/mma_numerics.cu -o tools/ubench/mma_numerics
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "tc_prims.cuh"

using namespace wetts::tc;

enum Mode { TF32_PLAIN = 0, TF32_X3 = 1, TF32_SPLIT2 = 2, F16_PLAIN = 3, F16_SPLIT2 = 4 };

struct NP {
  const float* A;   // [128][K]
  const float* B;   // [N][K]
  float* D;         // [128][N]
  int N, K, mode, nacc;
};

constexpr int KF = 64;   // k values per shared-memory fill

__global__ void __launch_bounds__(128, 1) k_numerics(NP p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool f16 = p.mode >= F16_PLAIN;
  const int es = f16 ? 2 : 4, kpc = 16 / es, groups = KF / kpc, N = p.N;
  const int kstep = f16 ? 16 : 8, steps = KF / kstep;
  uint8_t* Ahi = smem + 128;
  uint8_t* Alo = Ahi + groups * 2048;
  uint8_t* Bt = Alo + groups * 2048;                 // [groups][2N rows: hi | lo][16 B]
  if (warp == 0) tmem_alloc(smem_u32(slot), 512);
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init_fence(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t idesc_n = f16 ? idesc_f16_m128(N) : idesc_tf32_m128(N);
  const uint32_t idesc_2n = f16 ? idesc_f16_m128(2 * N) : idesc_tf32_m128(2 * N);
  const uint32_t bar = smem_u32(&bars[0]);
  uint32_t fill = 0, gstep = 0;
  for (int k0 = 0; k0 < p.K; k0 += KF, ++fill) {
    // ---- stage [128 x KF] of A and [N x KF] of B as hi / lo operand tiles
    for (int i = tid; i < 128 * KF; i += 128) {
      const int r = i / KF, k = i % KF;
      const float x = p.A[(size_t)r * p.K + k0 + k];
      const size_t off = (size_t)(k / kpc) * 2048 + (size_t)r * 16 + (size_t)(k % kpc) * es;
      if (f16) {
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn((x - __half2float(h)) * 2048.f);
        *reinterpret_cast<__half*>(Ahi + off) = h;
        *reinterpret_cast<__half*>(Alo + off) = l;
      } else {
        const float h = tf32_rna(x);
        *reinterpret_cast<float*>(Ahi + off) = h;
        *reinterpret_cast<float*>(Alo + off) = tf32_rna(x - h);
      }
    }
    for (int i = tid; i < N * KF; i += 128) {
      const int n = i / KF, k = i % KF;
      const float x = p.B[(size_t)n * p.K + k0 + k];
      const size_t g = (size_t)(k / kpc) * (2 * N) * 16;
      const size_t oh = g + (size_t)n * 16 + (size_t)(k % kpc) * es, ol = g + (size_t)(N + n) * 16 + (size_t)(k % kpc) * es;
      if (f16) {
        const __half h = __float2half_rn(x);
        *reinterpret_cast<__half*>(Bt + oh) = h;
        *reinterpret_cast<__half*>(Bt + ol) = __float2half_rn((x - __half2float(h)) * 2048.f);
      } else {
        const float h = tf32_rna(x);
        *reinterpret_cast<float*>(Bt + oh) = h;
        *reinterpret_cast<float*>(Bt + ol) = tf32_rna(x - h);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      for (int s = 0; s < steps; ++s, ++gstep) {
        const uint32_t acc = gstep % (uint32_t)p.nacc;
        const uint32_t first = (gstep < (uint32_t)p.nacc) ? 0u : 1u;   // first use of this accumulator set: overwrite
        const uint32_t d = tmem_base + acc * (uint32_t)(2 * N);
        const uint64_t a_hi = make_desc(smem_u32(Ahi) + (uint32_t)s * 2u * 2048u, 2048u, 128u);
        const uint64_t a_lo = make_desc(smem_u32(Alo) + (uint32_t)s * 2u * 2048u, 2048u, 128u);
        const uint32_t bbase = smem_u32(Bt) + (uint32_t)s * 2u * (uint32_t)(2 * N) * 16u;
        const uint64_t b_all = make_desc(bbase, (uint32_t)(2 * N) * 16u, 128u);               // rows 0..2N-1: [hi | lo]
        const uint64_t b_lo = make_desc(bbase + (uint32_t)N * 16u, (uint32_t)(2 * N) * 16u, 128u);
        switch (p.mode) {
          case TF32_PLAIN: tc_mma_tf32_1(d, a_hi, b_all, idesc_n, first); break;
          case F16_PLAIN: tc_mma_f16_1(d, a_hi, b_all, idesc_n, first); break;
          case TF32_X3:
            tc_mma_tf32_1(d, a_lo, b_all, idesc_n, first);
            tc_mma_tf32_1(d, a_hi, b_lo, idesc_n, 1u);
            tc_mma_tf32_1(d, a_hi, b_all, idesc_n, 1u);
            break;
          case TF32_SPLIT2:
            tc_mma_tf32_1(d, a_hi, b_all, idesc_2n, first);
            tc_mma_tf32_1(d + (uint32_t)N, a_lo, b_all, idesc_n, 1u);
            break;
          case F16_SPLIT2:
            tc_mma_f16_1(d, a_hi, b_all, idesc_2n, first);
            tc_mma_f16_1(d + (uint32_t)N, a_lo, b_all, idesc_n, 1u);
            break;
        }
      }
      if (elect_one()) tc_commit(bar);
      __syncwarp();
    }
    mbar_wait(bar, fill & 1);
    tc_fence_after();
    __syncthreads();
  }
  // ---- read back
  const bool split = (p.mode == TF32_SPLIT2 || p.mode == F16_SPLIT2);
  const float small_scale = (p.mode == F16_SPLIT2) ? (1.f / 2048.f) : 1.f;
  const int row = 32 * warp + lane;
  for (int c = 0; c < N; c += 16) {
    float out[16];
    for (int i = 0; i < 16; ++i) out[i] = 0.f;
    for (int a = 0; a < p.nacc; ++a) {
      float v[16], vs[16];
      tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(a * 2 * N + c), v);
      if (split) tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(a * 2 * N + N + c), vs);
      for (int i = 0; i < 16; ++i) out[i] += split ? (v[i] + vs[i] * small_scale) : v[i];
    }
    for (int i = 0; i < 16; ++i) p.D[(size_t)row * N + c + i] = out[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- timing: the fused kernels' pair pattern for both kinds (same 16 B-chunk addressing; f16 covers 16 channels per MMA)
struct P2 { int N, nblk, rp, iters, stride_rows, f16; long long* out; };
__global__ void __launch_bounds__(128, 1) k_pairs(P2 p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < (200 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(smem + 128)[i] = make_float4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(slot), 512);
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init_fence(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
  const uint32_t B_addr = smem_u32(smem + 128), A_addr = B_addr + 32 * 1024;
  const uint32_t a_half = (uint32_t)64 * p.rp * 4;
  const uint32_t idesc_n = p.f16 ? idesc_f16_m128(p.N) : idesc_tf32_m128(p.N);
  const uint32_t idesc_2n = p.f16 ? idesc_f16_m128(2 * p.N) : idesc_tf32_m128(2 * p.N);
  const uint64_t adesc0 = make_desc(A_addr, (uint32_t)p.rp * 16u, 128u);
  const uint64_t bdesc0 = make_desc(B_addr, (uint32_t)(2 * p.N) * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      const uint32_t tapshift = (uint32_t)(i & 7) * 3u;
      for (int m = 0; m < p.nblk; ++m) {
        const uint32_t a0 = alo0 + tapshift + (uint32_t)(m * p.stride_rows);
        const uint32_t d = tmem_base + (uint32_t)(m * 2 * p.N);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t al = a0 + (uint32_t)(kk * 2 * p.rp), bl = blo0 + (uint32_t)(kk * 2 * 2 * p.N);
          if (p.f16) {
            tc_mma_f16_1(d, desc_with_lo(adesc0, al), desc_with_lo(bdesc0, bl), idesc_2n, 1u);
            tc_mma_f16_1(d + (uint32_t)p.N, desc_with_lo(adesc0, al + (a_half >> 4)), desc_with_lo(bdesc0, bl), idesc_n, 1u);
          } else {
            tc_mma_tf32_split2(d, d + (uint32_t)p.N, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + (a_half >> 4)),
                               desc_with_lo(bdesc0, bl), idesc_2n, idesc_n, 1u);
          }
        }
      }
    }
    t1 = clock64();
    if (elect_one()) tc_commit(smem_u32(&bars[0]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    t2 = clock64();
    if (tid == 0) { p.out[blockIdx.x * 2] = t1 - t0; p.out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// plain MMA rate for wide N (per-layer kernel shape), both kinds
struct P3 { int N, iters, f16; long long* out; };
__global__ void __launch_bounds__(128, 1) k_wide(P3 p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < (200 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(smem + 128)[i] = make_float4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(slot), 512);
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init_fence(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
  const uint32_t A_addr = smem_u32(smem + 128), B_addr = A_addr + 100 * 1024;
  const uint32_t idesc = p.f16 ? idesc_f16_m128(p.N) : idesc_tf32_m128(p.N);
  const uint64_t adesc0 = make_desc(A_addr, 136u * 16u, 128u);
  const uint64_t bdesc0 = make_desc(B_addr, (uint32_t)p.N * 16u, 128u);
  long long t0 = 0, t2 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      const uint32_t al = (uint32_t)adesc0 + (uint32_t)(i & 3) * 2u + (uint32_t)((i >> 2) & 7) * 272u;
      const uint32_t bl = (uint32_t)bdesc0 + (uint32_t)((i >> 2) & 7) * (uint32_t)(2 * p.N);
      const uint32_t d = tmem_base + (uint32_t)((i & 1) * p.N);
      if (p.f16) tc_mma_f16_1(d, desc_with_lo(adesc0, al), desc_with_lo(bdesc0, bl), idesc, 1u);
      else tc_mma_tf32_1(d, desc_with_lo(adesc0, al), desc_with_lo(bdesc0, bl), idesc, 1u);
    }
    if (elect_one()) tc_commit(smem_u32(&bars[0]));
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    t2 = clock64();
    if (tid == 0) { p.out[blockIdx.x * 2] = t2 - t0; p.out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  // ------------------------------------------------------------ numerics
  const int N = 64;
  const char* names[] = {"tf32 plain (1 MMA)", "tf32 x3 (1 acc)", "tf32 split2", "f16 plain (1 MMA)", "f16 split2 scaled"};
  CK(cudaFuncSetAttribute(k_numerics, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("numerics: D[128 x %d] = A[128 x K] * B[%d x K]^T, A ~ N(0,1) (lrelu-like mix), B ~ N(0,1)/sqrt(K); fp64 reference\n", N, N);
  printf("%-22s %-6s %-5s | %12s %12s %12s\n", "mode", "K", "nacc", "max err/rms", "rms err/rms", "bias");
  for (int K : {192, 576, 2304}) {
    std::mt19937 rng(7 + K);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> A((size_t)128 * K), B((size_t)N * K), D((size_t)128 * N);
    for (auto& v : A) { v = nd(rng); if (v < 0) v *= 0.1f; }     // post-lrelu activations: positive bias like the real convs
    for (auto& v : B) v = nd(rng) / sqrtf((float)K);
    std::vector<double> ref((size_t)128 * N);
    double sq = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
        ref[(size_t)m * N + n] = s;
        sq += s * s;
      }
    const double rms = sqrt(sq / (128.0 * N));
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    for (int mode = 0; mode < 5; ++mode)
      for (int nacc : {1, 2, 4}) {
        if ((mode == TF32_PLAIN || mode == F16_PLAIN) && nacc > 1) continue;
        if (nacc * 2 * N > 512) continue;
        NP p{dA, dB, dD, N, K, mode, nacc};
        k_numerics<<<1, 128, 200 * 1024>>>(p);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double mx = 0, se = 0, dot = 0;
        for (size_t i = 0; i < D.size(); ++i) {
          const double e = (double)D[i] - ref[i];
          mx = fabs(e) > mx ? fabs(e) : mx;
          se += e * e;
          dot += e * ref[i];
        }
        printf("%-22s %-6d %-5d | %12.3e %12.3e %12.3e\n", names[mode], K, nacc, mx / rms, sqrt(se / D.size()) / rms, dot / sq);
      }
    // fp32 FMA reference error (sequential fp32 accumulation on the host) for scale
    {
      double mx = 0, se = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          float s = 0.f;
          for (int k = 0; k < K; ++k) s = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], s);
          const double e = (double)s - ref[(size_t)m * N + n];
          mx = fabs(e) > mx ? fabs(e) : mx;
          se += e * e;
        }
      printf("%-22s %-6d %-5s | %12.3e %12.3e\n", "host fp32 fma chain", K, "-", mx / rms, sqrt(se / (128.0 * N)) / rms);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  // ------------------------------------------------------------ timing
  long long* d_out;
  CK(cudaMalloc(&d_out, 1024 * 2 * sizeof(long long)));
  CK(cudaFuncSetAttribute(k_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  CK(cudaFuncSetAttribute(k_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  printf("\npair pattern (A_hi x [B_hi|B_lo] -> 2N cols, A_lo x B_hi -> N cols): cycles per PAIR; f16 pairs cover 16 channels, tf32 pairs 8\n");
  printf("%-6s %-5s %-6s %-6s | %10s %10s\n", "kind", "N", "nblk", "rp", "issue", "total");
  for (int f16 = 0; f16 < 2; ++f16)
    for (int Nn : {32, 64, 128})
      for (int nblk : {1, 2}) {
        if (nblk * 2 * Nn > 512) continue;
        P2 p{Nn, nblk, 225, 400, 72, f16, d_out};
        for (int rep = 0; rep < 2; ++rep) { k_pairs<<<148, 128, 210 * 1024>>>(p); CK(cudaDeviceSynchronize()); }
        std::vector<long long> h(148 * 2);
        CK(cudaMemcpy(h.data(), d_out, 148 * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
        double is = 0, tt = 0;
        for (int i = 0; i < 148; ++i) { is += h[2 * i]; tt += h[2 * i + 1]; }
        const double n = 148.0 * p.iters * nblk * 4;
        printf("%-6s %-5d %-6d %-6d | %10.1f %10.1f\n", f16 ? "f16" : "tf32", Nn, nblk, 225, is / n, tt / n);
      }
  printf("\nsingle MMAs, M = 128: cycles per MMA (tf32: K = 8, f16: K = 16)\n%-6s %-5s | %10s\n", "kind", "N", "total");
  for (int f16 = 0; f16 < 2; ++f16)
    for (int Nn : {64, 128, 192, 256}) {
      P3 p{Nn, 4000, f16, d_out};
      for (int rep = 0; rep < 2; ++rep) { k_wide<<<148, 128, 210 * 1024>>>(p); CK(cudaDeviceSynchronize()); }
      std::vector<long long> h(148 * 2);
      CK(cudaMemcpy(h.data(), d_out, 148 * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
      double tt = 0;
      for (int i = 0; i < 148; ++i) tt += h[2 * i];
      printf("%-6s %-5d | %10.1f\n", f16 ? "f16" : "tf32", Nn, tt / (148.0 * p.iters));
    }
  printf("done\n");
  return 0;
}
