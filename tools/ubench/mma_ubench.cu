// Micro-benchmark: sustained cycles per tcgen05.mma (kind::tf32, M = 128, SS operands, no-swizzle K-major)
// as a function of N, accumulator reuse, A-row start alignment and CTAs per SM.  One warp issues.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I wetts_b200/csrc tools/ubench/mma_ubench.cu -o tools/ubench/mma_ubench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_prims.cuh"

using namespace wetts::tc;

struct P {
  int N, n_acc, shift_mode, iters, x3, rp;
  long long* out;   // [grid][2]: issue cycles, total cycles
};

__global__ void __launch_bounds__(128, 1) k_ubench(P p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  // zero operands
  for (int i = tid; i < (200 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(smem + 128)[i] = make_float4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(slot), 512);
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init_fence(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
  const uint32_t A_addr = smem_u32(smem + 128), B_addr = A_addr + 120 * 1024;
  const uint32_t idesc = idesc_tf32_m128(p.N);
  const uint64_t adesc0 = make_desc(A_addr, (uint32_t)p.rp * 16u, 128u);
  const uint64_t bdesc0 = make_desc(B_addr, (uint32_t)p.N * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      uint32_t shift = 0;
      if (p.shift_mode == 1) shift = (uint32_t)(i & 7) * 3u;        // odd row starts (tap * dil)
      else if (p.shift_mode == 2) shift = (uint32_t)(i & 7) * 8u;   // 8-row aligned starts
      const uint32_t al = alo0 + shift, bl = blo0;
      const uint32_t d = tmem_base + (uint32_t)((i & (p.n_acc - 1)) * p.N);
      if (p.x3) {
        tc_mma_tf32_x3(d, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + 2048), desc_with_lo(bdesc0, bl),
                       desc_with_lo(bdesc0, bl + 1024), idesc, 1u);
      } else {
        asm volatile(
            "{\n\t.reg .pred pe;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;\n\t}" ::"r"(d),
            "l"(desc_with_lo(adesc0, al)), "l"(desc_with_lo(bdesc0, bl)), "r"(idesc)
            : "memory");
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

// two co-resident CTAs per SM: same kernel with 256 TMEM columns and ~100 KB smem
__global__ void __launch_bounds__(128, 2) k_ubench2(P p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<float4*>(smem + 128)[i] = make_float4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(smem_u32(slot), 256);
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init_fence(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  const uint32_t A_addr = smem_u32(smem + 128), B_addr = A_addr + 64 * 1024;
  const uint32_t idesc = idesc_tf32_m128(p.N);
  const uint64_t adesc0 = make_desc(A_addr, (uint32_t)p.rp * 16u, 128u);
  const uint64_t bdesc0 = make_desc(B_addr, (uint32_t)p.N * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      uint32_t shift = (p.shift_mode == 1) ? (uint32_t)(i % 7) * 3u : 0u;
      const uint32_t d = tmem_base + (uint32_t)((i % p.n_acc) * p.N);
      tc_mma_tf32_x3(d, desc_with_lo(adesc0, alo0 + shift), desc_with_lo(adesc0, alo0 + shift + 1024), desc_with_lo(bdesc0, blo0),
                     desc_with_lo(bdesc0, blo0 + 512), idesc, 1u);
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
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// The fused kernel's issue pattern: pairs (A_hi x [B_hi|B_lo] -> D[0:2N], A_lo x B_hi -> D[N:2N]); `nblk` accumulator
// blocks visited 4 pairs at a time (conv1: 2, conv2: 1); A row pitch rp (odd pitches misalign the second k-group).
struct P2 { int N, nblk, rp, iters, stride_rows; long long* out; };
__global__ void __launch_bounds__(128, 1) k_split2(P2 p) {
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
  const uint32_t B_addr = smem_u32(smem + 128), A_addr = B_addr + 32 * 1024;     // A: up to 2 x 64*225*4 = 115 KB
  const uint32_t a_half = (uint32_t)64 * p.rp * 4;
  const uint32_t idesc_n = idesc_tf32_m128(p.N), idesc_2n = idesc_tf32_m128(2 * p.N);
  const uint64_t adesc0 = make_desc(A_addr, (uint32_t)p.rp * 16u, 128u);
  const uint64_t bdesc0 = make_desc(B_addr, (uint32_t)(2 * p.N) * 16u, 128u);
  const uint32_t alo0 = (uint32_t)adesc0, blo0 = (uint32_t)bdesc0;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {          // one "chunk": nblk blocks x 4 pairs
      const uint32_t tapshift = (uint32_t)(i & 7) * 3u;
      for (int m = 0; m < p.nblk; ++m) {
        const uint32_t a0 = alo0 + tapshift + (uint32_t)(m * p.stride_rows);
        const uint32_t d = tmem_base + (uint32_t)(m * 2 * p.N);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t al = a0 + (uint32_t)(kk * 2 * p.rp), bl = blo0 + (uint32_t)(kk * 2 * 2 * p.N);
          tc_mma_tf32_split2(d, d + (uint32_t)p.N, desc_with_lo(adesc0, al), desc_with_lo(adesc0, al + (a_half >> 4)),
                             desc_with_lo(bdesc0, bl), idesc_2n, idesc_n, 1u);
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

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 1024 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k_ubench, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  cudaFuncSetAttribute(k_ubench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("%-8s %-6s %-6s %-6s %-4s %-6s | %12s %12s  (cycles per single MMA instruction)\n", "kernel", "N", "n_acc", "shift", "x3", "rp", "issue/mma", "total/mma");
  auto run = [&](int two, int N, int n_acc, int shift, int x3, int rp, int grid) {
    P p{N, n_acc, shift, 2000, x3, rp, d_out};
    for (int rep = 0; rep < 2; ++rep) {
      if (two) k_ubench2<<<grid, 128, 100 * 1024>>>(p);
      else k_ubench<<<grid, 128, 210 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    std::vector<long long> h(grid * 2);
    cudaMemcpy(h.data(), d_out, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    double is = 0, tt = 0;
    for (int i = 0; i < grid; ++i) { is += h[2 * i]; tt += h[2 * i + 1]; }
    const double n = (double)grid * p.iters * (x3 ? 3 : 1);
    printf("%-8s %-6d %-6d %-6d %-4d %-6d | %12.1f %12.1f\n", two ? "2cta/sm" : "1cta/sm", N, n_acc, shift, x3, rp, is / n, tt / n);
  };
  const int Ns[] = {32, 64, 128, 256};
  for (int N : Ns) run(0, N, 1, 0, 0, 224, 148);
  for (int N : Ns) run(0, N, 1, 0, 1, 224, 148);
  for (int N : Ns) if (N <= 128) run(0, N, 4, 0, 1, 224, 148);
  run(0, 32, 1, 1, 1, 224, 148);
  run(0, 32, 1, 2, 1, 224, 148);
  run(0, 64, 1, 1, 1, 224, 148);
  run(0, 64, 2, 1, 1, 224, 148);
  run(0, 32, 2, 1, 1, 128, 148);
  run(0, 32, 2, 1, 1, 136, 148);
  run(0, 32, 1, 0, 1, 224, 1);       // a single CTA on the whole chip
  run(1, 32, 1, 1, 1, 224, 296);
  run(1, 32, 2, 1, 1, 224, 296);
  run(1, 64, 1, 1, 1, 224, 296);
  cudaFuncSetAttribute(k_split2, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  printf("\nsplit2 pairs (fused kernel pattern): cycles per PAIR\n%-6s %-6s %-6s %-8s | %10s %10s\n", "N", "nblk", "rp", "stride", "issue", "total");
  auto run2 = [&](int N, int nblk, int rp, int stride) {
    P2 p{N, nblk, rp, 400, stride, d_out};
    for (int rep = 0; rep < 2; ++rep) {
      k_split2<<<148, 128, 210 * 1024>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    std::vector<long long> h(148 * 2);
    cudaMemcpy(h.data(), d_out, 148 * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    double is = 0, tt = 0;
    for (int i = 0; i < 148; ++i) { is += h[2 * i]; tt += h[2 * i + 1]; }
    const double n = 148.0 * p.iters * nblk * 4;
    printf("%-6d %-6d %-6d %-8d | %10.1f %10.1f\n", N, nblk, rp, stride, is / n, tt / n);
  };
  for (int N : {32, 64}) {
    run2(N, 1, 224, 0);
    run2(N, 1, 225, 0);
    run2(N, 2, 224, 72);
    run2(N, 2, 225, 72);
    run2(N, 2, 225, 128);
    run2(N, 2, 232, 72);
    run2(N, 2, 231, 72);
  }
  printf("done\n");
  return 0;
}
