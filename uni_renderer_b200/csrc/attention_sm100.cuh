// Fused attention for sm_100a -- host-visible parameter blocks (see attention_sm100.cu).
#pragma once
#include "common.cuh"

namespace unib {

struct AttnParams {
  int B, heads, Nq, Nk, d;
  float scale;       // softmax scale (d^-1/2)
  __half* out;       // [B*Nq, ldo] fp16; head h writes columns [h*d, (h+1)*d)
  int ldo;
  float* lse2;       // optional [B, heads, Nq]: log2-domain log-sum-exp of every row (training: the backward's input)
  int pdl_early;     // as GemmParams::pdl_early
  long long* trace;  // debug: clock64 stamps of CTA (0,0,0) (unib200_debug_set_trace); nullptr in production
};

struct alignas(64) AttnMaps {
  CUtensorMap q, k, v;   // 4-D {d, tokens, heads, batch}, box {64, 128 | BKV, 1, 1}, SWIZZLE_128B
};

int attention_bkv(int d);
cudaError_t launch_attention(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream);

}  // namespace unib
