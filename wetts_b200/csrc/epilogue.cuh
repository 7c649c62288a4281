// Conv epilogues shared by the fp32 SIMT kernel and the tcgen05 kernel, so both paths have
// bit-identical post-accumulation semantics (see EpiMode in kernels.cuh).
#pragma once
#include "kernels.cuh"

namespace wetts {

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// bias (+ per-(b,co) conditioning for EPI_PLAIN) of packed output channel `co`
__device__ __forceinline__ float channel_term(const ConvArgs& a, int b, int co) {
  float bv = a.bias ? a.bias[co] : 0.f;
  if (a.ep.mode == EPI_PLAIN && a.ep.cond) bv += a.ep.cond[(long long)b * a.ep.cond_bs + a.ep.cond_off + co];
  return bv;
}

// v already contains the bias term
__device__ __forceinline__ void epilogue_store(const ConvArgs& a, int b, int co, int t, float v, float msk) {
  const ConvEpilogue& e = a.ep;
  const int T = a.T;
  const long long off = (long long)b * e.out_bs + (long long)co * T + t;
  switch (e.mode) {
    case EPI_PLAIN:
      if (e.act == 1) v = fmaxf(v, 0.f);
      if (e.out_mask) v *= msk;
      e.out[off] = v;
      break;
    case EPI_RESID:
      e.out[off] = v + e.resid[off];
      break;
    case EPI_MRF: {
      v += e.resid[off];
      if (e.acc_mode == 0) e.out[off] = v;
      else if (e.acc_mode == 1) e.out[off] = e.out[off] + v;
      else e.out[off] = (e.out[off] + v) / e.div;
      break;
    }
    case EPI_RES_SKIP: {
      if (!e.last && co < e.H) {
        e.x[off] = (e.x[off] + v) * msk;
      } else {
        const int c2 = e.last ? co : co - e.H;
        const long long o2 = (long long)b * e.out_bs + (long long)c2 * T + t;
        e.skip[o2] = e.skip_init ? v : e.skip[o2] + v;
      }
      break;
    }
    case EPI_COUPLING: {
      const int zc = e.z_c0 + co * e.z_cstep;
      float* p = e.out + (long long)b * e.out_bs + (long long)zc * T + t;
      *p = (*p - v * msk) * msk;
      break;
    }
    default:
      break;
  }
}

// EPI_GATE: packed channels (co, co+1) = (tanh half j, sigmoid half j), j = co/2
__device__ __forceinline__ void gate_terms(const ConvArgs& a, int b, int co, float& ba, float& bb) {
  ba = 0.f;
  bb = 0.f;
  if (a.bias) { ba = a.bias[co]; bb = a.bias[co + 1]; }
  if (a.ep.cond) {
    const float* g = a.ep.cond + (long long)b * a.ep.cond_bs + a.ep.cond_off;
    ba += g[co >> 1];
    bb += g[a.ep.H + (co >> 1)];
  }
}
__device__ __forceinline__ void gate_store(const ConvArgs& a, int b, int co, int t, float va, float vb) {
  a.ep.out[(long long)b * a.ep.out_bs + (long long)(co >> 1) * a.T + t] = tanhf(va) * sigmoidf_acc(vb);
}

}  // namespace wetts
