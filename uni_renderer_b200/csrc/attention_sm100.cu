// FlashAttention-style fused softmax(Q K^T * scale) V for sm_100a (tcgen05 + TMEM + TMA), no mask.
//
// One CTA = 128 query rows of one (batch, head).  S = Q K^T and O += P V both run on tcgen05 with the accumulators
// in TMEM (S: BKV fp32 columns, O: round16(d) fp32 columns); the online softmax runs one thread per query row
// (tcgen05.ld 32x32b gives each thread its own row, so no cross-lane reductions), P is written back to shared
// memory as fp16 in the 128B-swizzled K-major layout and fed to the second MMA; O is rescaled in place in TMEM.
// Warp roles (192 threads): warps 0-3 softmax/epilogue, warp 4 TMA producer, warp 5 MMA issuer + TMEM owner.
// Q/K/V are read straight out of the (fused) projection outputs through 4-D tensor maps {d, tokens, heads, batch};
// head dims that are not multiples of 64 rely on TMA out-of-bounds zero fill, so nothing is padded in HBM.
#include "attention_sm100.cuh"

namespace unib {

template <int NCH, int BKV>
struct AttnCfg {
  static constexpr int kQBytes = NCH * 128 * 128;          // NCH chunks of [128 rows x 64 fp16]
  static constexpr int kKBytes = NCH * BKV * 128;          // NCH chunks of [BKV rows x 64 fp16]
  static constexpr int kStageBytes = 2 * kKBytes;          // K + V
  static constexpr int kPBytes = (BKV / 64) * 128 * 128;   // [128 rows x BKV fp16] as 64-wide chunks
  static constexpr int kStages = 2;
  static constexpr int kBarOff = kQBytes + kStages * kStageBytes + kPBytes;
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;
  static constexpr int kTmemCols = 256;                    // S (BKV <= 128) + O (<= 128 when BKV = 128, <= 192 when 64)
};

template <int NCH, int BKV>
__global__ void __launch_bounds__(192, 1)
attention_tcgen05_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NCH, BKV>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t q_smem = base;
  const uint32_t kv_smem = base + Cfg::kQBytes;
  const uint32_t p_smem = kv_smem + Cfg::kStages * Cfg::kStageBytes;
  const uint32_t bar_base = base + Cfg::kBarOff;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  const uint32_t s_full = bar_base + 8u * 5;
  const uint32_t p_full = bar_base + 8u * 6;
  const uint32_t o_ready = bar_base + 8u * 7;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kBarOff + 64);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (p.Nk + BKV - 1) / BKV;
  const int dpad = (p.d + 15) & ~15;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_ready, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Cfg::kQBytes);
      for (int ch = 0; ch < NCH; ++ch) tma_load_4d(q_smem + ch * 16384, &maps.q, q_full, ch * 64, q0, head, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(kv_empty(st), ph ^ 1);
        mbar_arrive_expect_tx(kv_full(st), Cfg::kStageBytes);
        const uint32_t kdst = kv_smem + st * Cfg::kStageBytes;
        for (int ch = 0; ch < NCH; ++ch) {
          tma_load_4d(kdst + ch * (BKV * 128), &maps.k, kv_full(st), ch * 64, j * BKV, head, b);
          tma_load_4d(kdst + Cfg::kKBytes + ch * (BKV * 128), &maps.v, kv_full(st), ch * 64, j * BKV, head, b);
        }
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, BKV);
      const uint32_t idesc_o = make_idesc_f16(128, dpad, 0, 1);   // B (= V) is MN-major
      const uint32_t t_s = tmem_base;
      const uint32_t t_o = tmem_base + BKV;
      mbar_wait(q_full, 0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(kv_full(st), ph);
        tc_fence_after();
        const uint32_t k_addr = kv_smem + st * Cfg::kStageBytes;
        const uint32_t v_addr = k_addr + Cfg::kKBytes;
        // S = Q K^T : K loop over the head dim in steps of 16
        for (int ks = 0; ks < dpad / 16; ++ks) {
          const int ch = ks >> 2, within = ks & 3;
          const uint64_t a_desc = make_desc_kmajor_sw128(q_smem + ch * 16384 + within * 32);
          const uint64_t b_desc = make_desc_kmajor_sw128(k_addr + ch * (BKV * 128) + within * 32);
          umma_f16_ss(t_s, a_desc, b_desc, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        // O += P V : K loop over the kv rows of this block in steps of 16
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const int ch = ks >> 2, within = ks & 3;
          const uint64_t a_desc = make_desc_kmajor_sw128(p_smem + ch * 16384 + within * 32);
          const uint64_t b_desc = make_desc_mnmajor_sw128(v_addr + ks * 2048, BKV * 128, 1024);
          umma_f16_ss(t_o, a_desc, b_desc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(kv_empty(st));
        umma_commit(o_ready);
      }
    }
  } else {
    // =============================== softmax + epilogue (warps 0..3) ===============================
    const int qd = warp;                    // TMEM lane quadrant
    const int row = qd * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const uint32_t t_o = t_s + BKV;
    const float sl2 = p.scale * 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t p_row = p_smem + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = p.Nk - j * BKV;     // columns >= kv_valid are padding
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        float v[32];
        tmem_ld32(t_s + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (c * 32 + i < kv_valid) ? v[i] : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2((m_run - m_new) * sl2);
      const float mb = m_new * sl2;
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV / 32; ++c) {
        float v[32];
        tmem_ld32(t_s + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = fast_exp2(v[i] * sl2 - mb);
          float p1 = fast_exp2(v[i + 1] * sl2 - mb);
          if (c * 32 + i >= kv_valid) p0 = 0.f;
          if (c * 32 + i + 1 >= kv_valid) p1 = 0.f;
          rowsum += p0 + p1;
          pk[i >> 1] = pack_half2(p0, p1);
        }
        // 32 columns = 4 x 16 B units; unit u of the 64-wide chunk lands at (u ^ (row & 7))
        const int chunk = (c * 32) >> 6;
        const int u0 = ((c * 32) & 63) >> 3;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t dst = p_row + chunk * 16384 + (((u0 + u) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * u]), "r"(pk[4 * u + 1]),
                       "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3])
                       : "memory");
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      if (j > 0) {
        mbar_wait(o_ready, (j - 1) & 1);      // previous P V must have landed before O is rescaled
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < dpad / 16; ++c) {
          float o[16];
          tmem_ld16(t_o + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] *= alpha;
          tmem_st16(t_o + c * 16, o);
        }
        tmem_st_wait();
      }
      fence_proxy_async_shared();             // P (generic-proxy stores) -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_ready, (nblk - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int q = q0 + row;
    __half* op = p.out + (static_cast<size_t>(b) * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll 1
    for (int c = 0; c < dpad / 16; ++c) {
      float o[16];
      tmem_ld16(t_o + c * 16, o);
      tmem_ld_wait();
      if (q < p.Nq) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = c * 16 + u * 8;
          if (col + 8 <= p.d) {
            uint4 w;
            w.x = pack_half2(o[u * 8 + 0] * inv_l, o[u * 8 + 1] * inv_l);
            w.y = pack_half2(o[u * 8 + 2] * inv_l, o[u * 8 + 3] * inv_l);
            w.z = pack_half2(o[u * 8 + 4] * inv_l, o[u * 8 + 5] * inv_l);
            w.w = pack_half2(o[u * 8 + 6] * inv_l, o[u * 8 + 7] * inv_l);
            *reinterpret_cast<uint4*>(op + col) = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int NCH, int BKV>
static cudaError_t launch_cfg(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream) {
  using Cfg = AttnCfg<NCH, BKV>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tcgen05_kernel<NCH, BKV>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.Nq + 127) / 128, p.heads, p.B);
  attention_tcgen05_kernel<NCH, BKV><<<grid, 192, Cfg::kSmemBytes, stream>>>(maps, p);
  return cudaGetLastError();
}

int attention_bkv(int d) { return d <= 128 ? 128 : 64; }

cudaError_t launch_attention(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream) {
  if (p.d % 8 != 0 || p.d > 192 || p.d < 8) return cudaErrorInvalidValue;
  if (p.d <= 64) return launch_cfg<1, 128>(maps, p, stream);
  if (p.d <= 128) return launch_cfg<2, 128>(maps, p, stream);
  return launch_cfg<3, 64>(maps, p, stream);
}

}  // namespace unib
