// Flash-attention backward for sm_100a -- host-visible parameter blocks (see attention_bwd_sm100.cu).
#pragma once
#include "common.cuh"

namespace unib {

struct AttnBwdParams {
  int B, heads, Nq, Nk, d;
  float scale;            // softmax scale (d^-1/2)
  const float* lse2;      // [B, heads, Nq] log2-domain log-sum-exp of the forward (unib200_attn_desc.lse2)
  float* D;               // [B, heads, Nq] scratch: rowsum(dO o O), filled by the prep kernel
  float* dq_acc;          // [B*Nq, ld_dq] fp32 accumulator (zeroed by the caller); head h adds to columns [h*d, (h+1)*d)
  int ld_dq;
  __half* dk; int ld_dk;  // [B*Nk, ld] fp16 outputs, head h writes columns [h*d, (h+1)*d)
  __half* dv; int ld_dv;
};

struct alignas(64) AttnBwdMaps {
  CUtensorMap q, k, v, dout;   // 4-D {d, tokens, heads, batch}, box {64, 128, 1, 1}, SWIZZLE_128B
};

cudaError_t launch_attention_bwd(const AttnBwdMaps& maps, const AttnBwdParams& p, const __half* o, int ldo,
                                 const __half* dout, int lddo, cudaStream_t stream);

}  // namespace unib
