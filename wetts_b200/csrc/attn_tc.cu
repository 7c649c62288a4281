// Launch side of the tensor-pipe attention kernel (attn_tc_kernel.cuh).
#include <cstdlib>

#include "attn_tc_kernel.cuh"
#include "kernels.cuh"

namespace wetts {

bool rel_attention_tc_supported(int C, int T, int n_heads, int window) {
  return n_heads > 0 && C == n_heads * kAttnTcDk && T >= 64 && T <= kAttnTcMaxT && window == 4;
}

int launch_rel_attention_tc(const float* qkv, const float* emb_k, const float* emb_v, const long long* lengths, float* out,
                            int B, int C, int T, int n_heads, int window, cudaStream_t s) {
  AttnTcArgs a;
  a.qkv = qkv; a.emb_k = emb_k; a.emb_v = emb_v; a.lengths = lengths; a.out = out;
  a.B = B; a.C = C; a.T = T; a.n_heads = n_heads; a.window = window;
  if (dyn_smem_offset(&a.smem_off, s)) return 1;
  // two CTAs per SM (shared memory used twice, 256 TMEM columns): default since measured (240 -> 149 us per launch at the
  // bench shape, profiles/r03d_*; Tx >= 64 fixtures green with it); WETTS_ATTN_TC_CTAS=1 = the
  // one-CTA layout (208 KB, 512 columns)
  static const int ctas = getenv("WETTS_ATTN_TC_CTAS") ? atoi(getenv("WETTS_ATTN_TC_CTAS")) : 2;
  if (ctas >= 2) {
    static DynSmemAttr attr2;
    if (attr2.ensure((const void*)rel_attention_tc_kernel<true>, kAttnTcSmemShared) != cudaSuccess) return 1;
    rel_attention_tc_kernel<true><<<B * n_heads, kAttnTcThreads, kAttnTcSmemShared, s>>>(a);
  } else {
    static DynSmemAttr attr;
    if (attr.ensure((const void*)rel_attention_tc_kernel<false>, kAttnTcSmem) != cudaSuccess) return 1;
    rel_attention_tc_kernel<false><<<B * n_heads, kAttnTcThreads, kAttnTcSmem, s>>>(a);
  }
  count_launch();
  return 0;
}

int attn_tc_install_fault_word(unsigned int* word) { return tc::install_fault_word_tu(word) == cudaSuccess ? 0 : 1; }

}  // namespace wetts
