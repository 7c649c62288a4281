// fp32 SIMT convolution kernels for the VITS hot path (sm_100a).
//
// conv1d_kernel: dilated Conv1d as an implicit GEMM.  One CTA computes a
// [CO_TILE x T_TILE] output tile; the input tile (with halo, pre-activation and
// masks applied once at staging) and a [CI_CHUNK x K x CO_TILE] weight tile live in
// shared memory; each thread owns an 8(co) x TN(t) register tile with the t's strided
// by 32 so that every shared-memory read of activations is a conflict-free 128B
// wavefront for any dilation, and weight reads are warp-uniform float4 broadcasts.
// Epilogues fuse bias, residual adds, the MRF mean, the WaveNet gate, the res/skip
// split and the coupling update, so no elementwise kernel touches HBM in between.
#include <atomic>
#include <cstdio>
#include <mutex>

#include "kernels.cuh"
#include "epilogue.cuh"

namespace wetts {

static std::atomic<unsigned long long> g_launch_count{0};
unsigned long long kernel_launch_counter() { return g_launch_count.load(std::memory_order_relaxed); }
void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

static thread_local cudaError_t g_launcher_error = cudaSuccess;
void note_launcher_error(cudaError_t e) { if (e != cudaSuccess && g_launcher_error == cudaSuccess) g_launcher_error = e; }
cudaError_t take_launcher_error() { cudaError_t e = g_launcher_error; g_launcher_error = cudaSuccess; return e; }

cudaError_t DynSmemAttr::ensure(const void* func, size_t bytes) {
  static std::mutex mu_all;   // growth is rare (a handful of times per process): one lock for every launch site
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { note_launcher_error(e); return e; }
  if (dev < 0 || dev >= kMaxDev) {   // uncached device index: set it every time
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    note_launcher_error(e);
    return e;
  }
  if (bytes <= __atomic_load_n(&have[dev], __ATOMIC_ACQUIRE)) return cudaSuccess;
  std::lock_guard<std::mutex> lock(mu_all);
  if (bytes <= have[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) __atomic_store_n(&have[dev], bytes, __ATOMIC_RELEASE);
  note_launcher_error(e);
  return e;
}

int current_device_sm_count() {
  static std::atomic<int> cache[DynSmemAttr::kMaxDev];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < DynSmemAttr::kMaxDev) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (dev >= 0 && dev < DynSmemAttr::kMaxDev) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

namespace {

constexpr int kThreads = 256;
constexpr int kCiChunk = 16;
constexpr int kTM = 8;

template <int TN, int WARPS_CO>
__global__ void __launch_bounds__(kThreads, 2) conv1d_kernel(const ConvArgs a) {
  constexpr int WARPS_T = 8 / WARPS_CO;
  constexpr int CO_TILE = kTM * WARPS_CO;
  constexpr int T_TILE = 32 * TN * WARPS_T;
  extern __shared__ __align__(16) float smem[];
  const int K = a.K, dil = a.dil;
  const int XS = T_TILE + (K - 1) * dil;
  float* xs = smem;                  // [kCiChunk][XS]
  float* ws = smem + kCiChunk * XS;  // [kCiChunk][K][CO_TILE]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_co = warp % WARPS_CO, warp_t = warp / WARPS_CO;
  const int b = blockIdx.z;
  const int co_blk = blockIdx.y * CO_TILE;
  const int t0 = blockIdx.x * T_TILE;
  const int co0 = warp_co * kTM;
  const int tl = warp_t * 32 * TN + lane;
  const int T = a.T;
  if (a.la_len && (long long)t0 >= (a.la_len[b] + a.la_margin) * (long long)a.la_rate) return;   // length-aware: whole tile unused
  const long long len = a.lengths ? a.lengths[b] : (long long)T;

  float acc[kTM][TN];
#pragma unroll
  for (int m = 0; m < kTM; ++m)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[m][j] = 0.f;

  const float* in_b = a.in + (long long)b * a.in_bs;
  const int t_in0 = t0 - a.pad_left;
  const int t_hi = a.in_mask ? (int)(len < T ? len : T) : T;

  for (int c0 = 0; c0 < a.Cin; c0 += kCiChunk) {
    __syncthreads();
    // ---- stage activations (pre-activation + mask + zero padding folded in)
    for (int ci = warp; ci < kCiChunk; ci += kThreads / 32) {
      const bool cok = (c0 + ci) < a.Cin;
      const float* src = in_b + (long long)(c0 + ci) * a.in_cs;
      float* dst = xs + ci * XS;
      for (int tt = lane; tt < XS; tt += 32) {
        const int t = t_in0 + tt;
        float v = 0.f;
        if (cok && t >= 0 && t < t_hi) {
          v = __ldg(src + t);
          if (a.pre_act) v = v > 0.f ? v : v * a.pre_slope;
        }
        dst[tt] = v;
      }
    }
    // ---- stage weights: rows (ci,k) are contiguous in the packed layout
    {
      constexpr int C4 = CO_TILE / 4;
      const int rows = kCiChunk * K;
      const int row_lim = (a.Cin - c0) * K;
      const float* wsrc = a.w + (long long)c0 * K * a.CoutPad + co_blk;
      for (int idx = tid; idx < rows * C4; idx += kThreads) {
        const int row = idx / C4, c4 = idx - row * C4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_lim) v = __ldg(reinterpret_cast<const float4*>(wsrc + (long long)row * a.CoutPad) + c4);
        reinterpret_cast<float4*>(ws + row * CO_TILE)[c4] = v;
      }
    }
    __syncthreads();
    // ---- multiply-accumulate
    const int nci = min(kCiChunk, a.Cin - c0);
    for (int ci = 0; ci < nci; ++ci) {
      const float* xr = xs + ci * XS + tl;
      const float* wr = ws + ci * K * CO_TILE + co0;
#pragma unroll 1
      for (int k = 0; k < K; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(wr);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
        wr += CO_TILE;
        float xv[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) xv[j] = xr[32 * j];
        xr += dil;
        const float wv[kTM] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int m = 0; m < kTM; ++m)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[m][j] = fmaf(wv[m], xv[j], acc[m][j]);
      }
    }
  }

  // ---- epilogue (shared with the tensor-core kernel: epilogue.cuh)
  const int co_base = co_blk + co0;
  if (a.ep.mode == EPI_GATE) {
#pragma unroll
    for (int m = 0; m < kTM; m += 2) {
      const int co = co_base + m;
      if (co >= a.Cout) continue;
      float ba, bb;
      gate_terms(a, b, co, ba, bb);
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int t = t0 + tl + 32 * j;
        if (t < T) gate_store(a, b, co, t, acc[m][j] + ba, acc[m + 1][j] + bb);
      }
    }
    return;
  }
#pragma unroll
  for (int m = 0; m < kTM; ++m) {
    const int co = co_base + m;
    if (co >= a.Cout) continue;
    const float bv = channel_term(a, b, co);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int t = t0 + tl + 32 * j;
      if (t < T) epilogue_store(a, b, co, t, acc[m][j] + bv, (t < len) ? 1.f : 0.f);
    }
  }
}

template <int TN, int WARPS_CO>
void launch_conv_inst(const ConvArgs& a, cudaStream_t s) {
  constexpr int WARPS_T = 8 / WARPS_CO;
  constexpr int CO_TILE = kTM * WARPS_CO;
  constexpr int T_TILE = 32 * TN * WARPS_T;
  const int XS = T_TILE + (a.K - 1) * a.dil;
  const size_t smem = sizeof(float) * ((size_t)kCiChunk * XS + (size_t)kCiChunk * a.K * CO_TILE);
  static DynSmemAttr attr;
  if (attr.ensure((const void*)conv1d_kernel<TN, WARPS_CO>, smem) != cudaSuccess) return;
  dim3 grid((a.T + T_TILE - 1) / T_TILE, a.CoutPad / CO_TILE, a.B);
  conv1d_kernel<TN, WARPS_CO><<<grid, kThreads, smem, s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------ ConvTranspose1d (polyphase)
constexpr int kCiChunkT = 8;

template <int TN, int WARPS_CO>
__global__ void __launch_bounds__(kThreads, 2) convT_kernel(const ConvTArgs a) {
  constexpr int WARPS_T = 8 / WARPS_CO;
  constexpr int CO_TILE = kTM * WARPS_CO;
  constexpr int N_TILE = 32 * TN * WARPS_T;
  extern __shared__ __align__(16) float smem[];
  const int u = a.u, ntaps = a.ntaps;
  const int QN = N_TILE / u + ntaps;
  float* ws = smem;                                      // [kCiChunkT][ntaps][CO_TILE][u]
  float* xs = smem + kCiChunkT * ntaps * CO_TILE * u;    // [kCiChunkT][QN]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int warp_co = warp % WARPS_CO, warp_t = warp / WARPS_CO;
  const int b = blockIdx.z;
  const int co_blk = blockIdx.y * CO_TILE;
  const int n0 = blockIdx.x * N_TILE;
  const int co0 = warp_co * kTM;
  const int nl = warp_t * 32 * TN + lane;
  const int Tout = a.T * u;
  const int q_lo = (n0 + a.pad) / u - (ntaps - 1);
  const int r = (n0 + nl + a.pad) % u;          // same for every j because 32 % u == 0
  const int qq0 = (n0 + nl + a.pad) / u - q_lo;  // local input index of j = 0
  const int qstep = 32 / u;

  float acc[kTM][TN];
#pragma unroll
  for (int m = 0; m < kTM; ++m)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[m][j] = 0.f;

  const float* in_b = a.in + (long long)b * a.Cin * a.T;
  for (int c0 = 0; c0 < a.Cin; c0 += kCiChunkT) {
    __syncthreads();
    for (int idx = tid; idx < kCiChunkT * QN; idx += kThreads) {
      const int ci = idx / QN, qq = idx - ci * QN;
      const int q = q_lo + qq;
      float v = 0.f;
      if (c0 + ci < a.Cin && q >= 0 && q < a.T) {
        v = __ldg(in_b + (long long)(c0 + ci) * a.T + q);
        v = v > 0.f ? v : v * a.pre_slope;
      }
      xs[idx] = v;
    }
    {
      const int rowlen = CO_TILE * u;  // contiguous in the packed layout
      const int R4 = rowlen / 4;
      const int rows = kCiChunkT * ntaps;
      const int row_lim = (a.Cin - c0) * ntaps;
      const float* wsrc = a.w + ((long long)c0 * ntaps * a.CoutPad + co_blk) * u;
      for (int idx = tid; idx < rows * R4; idx += kThreads) {
        const int row = idx / R4, c4 = idx - row * R4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_lim) v = __ldg(reinterpret_cast<const float4*>(wsrc + (long long)row * a.CoutPad * u) + c4);
        reinterpret_cast<float4*>(ws + row * rowlen)[c4] = v;
      }
    }
    __syncthreads();
    const int nci = min(kCiChunkT, a.Cin - c0);
    for (int ci = 0; ci < nci; ++ci) {
#pragma unroll 1
      for (int tap = 0; tap < ntaps; ++tap) {
        const float* wr = ws + ((ci * ntaps + tap) * CO_TILE + co0) * u + r;
        const float* xr = xs + ci * QN + qq0 - tap;
        float wv[kTM], xv[TN];
#pragma unroll
        for (int m = 0; m < kTM; ++m) wv[m] = wr[m * u];
#pragma unroll
        for (int j = 0; j < TN; ++j) xv[j] = xr[j * qstep];
#pragma unroll
        for (int m = 0; m < kTM; ++m)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[m][j] = fmaf(wv[m], xv[j], acc[m][j]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < kTM; ++m) {
    const int co = co_blk + co0 + m;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? a.bias[co] : 0.f;
    float* o = a.out + ((long long)b * a.Cout + co) * Tout;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + nl + 32 * j;
      if (n < Tout) o[n] = acc[m][j] + bv;
    }
  }
}

template <int TN, int WARPS_CO>
void launch_convT_inst(const ConvTArgs& a, cudaStream_t s) {
  constexpr int WARPS_T = 8 / WARPS_CO;
  constexpr int CO_TILE = kTM * WARPS_CO;
  constexpr int N_TILE = 32 * TN * WARPS_T;
  const int QN = N_TILE / a.u + a.ntaps;
  const size_t smem = sizeof(float) * ((size_t)kCiChunkT * a.ntaps * CO_TILE * a.u + (size_t)kCiChunkT * QN);
  static DynSmemAttr attr;
  if (attr.ensure((const void*)convT_kernel<TN, WARPS_CO>, smem) != cudaSuccess) return;
  dim3 grid((a.T * a.u + N_TILE - 1) / N_TILE, a.CoutPad / CO_TILE, a.B);
  convT_kernel<TN, WARPS_CO><<<grid, kThreads, smem, s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------ conv_post + tanh
constexpr int kPostTile = 1024;
constexpr int kPostCi = 8;

__global__ void __launch_bounds__(kThreads) conv_post_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                            float* __restrict__ out, int C, int T, int K, float slope) {
  extern __shared__ __align__(16) float smem[];
  const int XS = kPostTile + K - 1;
  float* xs = smem;                 // [kPostCi][XS]
  float* wsm = smem + kPostCi * XS;  // [C][K]
  const int b = blockIdx.y, t0 = blockIdx.x * kPostTile, tid = threadIdx.x;
  const int pad = (K - 1) / 2;
  for (int i = tid; i < C * K; i += kThreads) wsm[i] = w[i];
  float acc[kPostTile / kThreads];
#pragma unroll
  for (int j = 0; j < kPostTile / kThreads; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < C; c0 += kPostCi) {
    __syncthreads();
    for (int idx = tid; idx < kPostCi * XS; idx += kThreads) {
      const int ci = idx / XS, tt = idx - ci * XS;
      const int t = t0 - pad + tt;
      float v = 0.f;
      if (c0 + ci < C && t >= 0 && t < T) {
        v = __ldg(in + ((long long)b * C + c0 + ci) * T + t);
        v = v > 0.f ? v : v * slope;
      }
      xs[idx] = v;
    }
    __syncthreads();
    const int nci = min(kPostCi, C - c0);
    for (int ci = 0; ci < nci; ++ci) {
      for (int k = 0; k < K; ++k) {
        const float wv = wsm[(c0 + ci) * K + k];
#pragma unroll
        for (int j = 0; j < kPostTile / kThreads; ++j) acc[j] = fmaf(wv, xs[ci * XS + tid + j * kThreads + k], acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kPostTile / kThreads; ++j) {
    const int t = t0 + tid + j * kThreads;
    if (t < T) out[(long long)b * T + t] = tanhf(acc[j]);
  }
}

// Same op for K = 7, T % 4 == 0: every thread produces 4 consecutive samples from three 16 B loads per
// channel (its own quad plus the two neighbouring quads, which its neighbour threads also load: L1 hits), so
// the kernel is a pure streaming read of the activation (HBM-bound) instead of one shared-memory load per FMA.
__global__ void __launch_bounds__(kThreads) conv_post7_vec_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                 float* __restrict__ out, int C, int T, float slope,
                                                                 const long long* __restrict__ la_len, int la_rate, int la_margin) {
  extern __shared__ __align__(16) float wsm[];   // [C][8] (tap 7 = 0)
  const int b = blockIdx.y, tid = threadIdx.x;
  if (la_len && (long long)blockIdx.x * kThreads * 4 >= (la_len[b] + la_margin) * (long long)la_rate) return;   // length-aware
  for (int i = tid; i < C * 8; i += kThreads) wsm[i] = ((i & 7) < 7) ? w[(i >> 3) * 7 + (i & 7)] : 0.f;
  __syncthreads();
  const int q = blockIdx.x * kThreads + tid;   // quad index
  const int t4 = 4 * q;
  if (t4 >= T) return;
  const float* xb = in + (long long)b * C * T + t4;
  const bool has_l = t4 >= 4, has_r = t4 + 8 <= T;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* xc = xb + (long long)c * T;
    const float4 l = has_l ? __ldg(reinterpret_cast<const float4*>(xc - 4)) : z4;
    const float4 m = __ldg(reinterpret_cast<const float4*>(xc));
    const float4 r = has_r ? __ldg(reinterpret_cast<const float4*>(xc + 4)) : z4;
    float x[10] = {l.y, l.z, l.w, m.x, m.y, m.z, m.w, r.x, r.y, r.z};   // samples t4-3 .. t4+6
#pragma unroll
    for (int i = 0; i < 10; ++i) x[i] = x[i] > 0.f ? x[i] : x[i] * slope;
    const float4 w0 = *reinterpret_cast<const float4*>(wsm + 8 * c);
    const float4 w1 = *reinterpret_cast<const float4*>(wsm + 8 * c + 4);
    const float wk[7] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      acc0 = fmaf(wk[k], x[k + 0], acc0);
      acc1 = fmaf(wk[k], x[k + 1], acc1);
      acc2 = fmaf(wk[k], x[k + 2], acc2);
      acc3 = fmaf(wk[k], x[k + 3], acc3);
    }
  }
  *reinterpret_cast<float4*>(out + (long long)b * T + t4) = make_float4(tanhf(acc0), tanhf(acc1), tanhf(acc2), tanhf(acc3));
}

// Per-utterance conditioning vectors out[b][co] = bias[co] + sum_ci W[co][ci] g[b][ci] (a 1x1 conv over T = 1,
// i.e. a small GEMM over the batch).  Block = 64 output channels x 16 utterances; w is the SIMT layout
// [Cin][1][CoutPad] (output channel contiguous: coalesced), g is staged in shared memory.
constexpr int kCondB = 16;
__global__ void __launch_bounds__(256) cond_vec_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out, int B,
                                                       int Cin, int Cout, int CoutPad) {
  extern __shared__ float gs[];   // [kCondB][Cin]
  const int co = blockIdx.x * 64 + (threadIdx.x & 63), bq = threadIdx.x >> 6, b0 = blockIdx.y * kCondB;
  for (int i = threadIdx.x; i < kCondB * Cin; i += 256) {
    const int b = b0 + i / Cin;
    gs[i] = b < B ? g[(long long)b * Cin + (i % Cin)] : 0.f;
  }
  __syncthreads();
  if (co >= Cout) return;
  float acc[4];
  const float bv = bias ? bias[co] : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = bv;
  for (int ci = 0; ci < Cin; ++ci) {
    const float wv = __ldg(w + (long long)ci * CoutPad + co);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv, gs[(bq * 4 + j) * Cin + ci], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = b0 + bq * 4 + j;
    if (b < B) out[(long long)b * Cout + co] = acc[j];
  }
}

}  // namespace

void launch_cond_vector(const float* g, const float* w, const float* bias, float* out, int B, int Cin, int Cout,
                        int CoutPad, cudaStream_t s) {
  dim3 grid((Cout + 63) / 64, (B + kCondB - 1) / kCondB);
  cond_vec_kernel<<<grid, 256, (size_t)kCondB * Cin * sizeof(float), s>>>(g, w, bias, out, B, Cin, Cout, CoutPad);
  count_launch();
}

void launch_conv1d(const ConvArgs& a, cudaStream_t s) {
  if (a.wtc16 && a.fmt == 16 && a.T >= 64 && a.dil == a.tc16.dil && a.use_tc) launch_conv1d_tc16(a, s);
  else if (a.wtc && a.T >= 64 && a.dil == a.tc.dil && a.use_tc) launch_conv1d_tc(a, s);
  else launch_conv1d_simt(a, s);
}

void launch_conv1d_simt(const ConvArgs& a, cudaStream_t s) {
  const bool wide = (a.CoutPad % 64) == 0;
  const int T = a.T;
  if (wide) {
    if (T <= 32) launch_conv_inst<1, 8>(a, s);
    else if (T <= 64) launch_conv_inst<2, 8>(a, s);
    else if (T <= 160) launch_conv_inst<4, 8>(a, s);
    else launch_conv_inst<8, 8>(a, s);
  } else {
    if (T <= 64) launch_conv_inst<1, 4>(a, s);
    else if (T <= 128) launch_conv_inst<2, 4>(a, s);
    else if (T <= 320) launch_conv_inst<4, 4>(a, s);
    else launch_conv_inst<8, 4>(a, s);
  }
}

void launch_conv_transpose1d(const ConvTArgs& a, cudaStream_t s) {
  const bool wide = (a.CoutPad % 64) == 0;
  const int N = a.T * a.u;
  if (wide) {
    if (N <= 160) launch_convT_inst<4, 8>(a, s);
    else launch_convT_inst<8, 8>(a, s);
  } else {
    if (N <= 320) launch_convT_inst<4, 4>(a, s);
    else launch_convT_inst<8, 4>(a, s);
  }
}

void launch_conv_post_tanh(const float* in, const float* w, float* out, int B, int C, int T, int K, float slope,
                           cudaStream_t s, const long long* la_len, int la_rate, int la_margin) {
  if (K == 7 && (T & 3) == 0 && C <= 512 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
    dim3 grid((T / 4 + kThreads - 1) / kThreads, B);
    conv_post7_vec_kernel<<<grid, kThreads, (size_t)C * 8 * sizeof(float), s>>>(in, w, out, C, T, slope, la_len, la_rate, la_margin);
    count_launch();
    return;
  }
  const size_t smem = sizeof(float) * ((size_t)kPostCi * (kPostTile + K - 1) + (size_t)C * K);
  static DynSmemAttr attr;
  if (attr.ensure((const void*)conv_post_kernel, smem) != cudaSuccess) return;
  dim3 grid((T + kPostTile - 1) / kPostTile, B);
  conv_post_kernel<<<grid, kThreads, smem, s>>>(in, w, out, C, T, K, slope);
  count_launch();
}

}  // namespace wetts
