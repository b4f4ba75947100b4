// Flash-attention BACKWARD for sm_100a (tcgen05 + TMEM + TMA), head dims <= 80, no mask -- the gradient of
// O = softmax(Q K^T * scale) V used by the training path (uni_renderer_b200/trainer.py; the reference trains through
// diffusers' AttnProcessor2_0 = F.scaled_dot_product_attention, train/train.py:1421 accelerator.backward).
//
// One CTA owns ONE 128-key block of one (batch, head) and walks all 128-row query tiles.  Per query tile i:
//     S  = Q_i K^T,  dP = dO_i V^T                    two tcgen05 GEMMs into TMEM (128 x 128 fp32 each)
//     P  = exp2(S * scale*log2e - lse2_i)             softmax warps, one thread per query row; lse2 = the forward
//     dS = P o (dP - D_i) * scale                     kernel's log2-domain log-sum-exp, D_i = rowsum(dO_i o O_i)
//     P, dS -> shared memory as fp16 [q][k] in the 128-byte-swizzled 64-column chunks the forward uses for P
//     dV += P^T dO_i,  dK += dS^T Q_i                 the SAME [q][k] tiles read as MN-major ("transposed") A operands
//     dQ_i  = dS K                                    K-major A; K block as MN-major B (like V in the forward's P V)
//     dQ_i -> fp32 atomics into the dQ accumulator    (every key block contributes; converted to fp16 afterwards)
// dK / dV of the block stay in TMEM for the whole walk and are written once.  TMEM: S 128 | dP 128 | dV, dK, dQ
// round16(d) columns each (hence d <= 80).  The walk is NOT software-pipelined (phases run back to back behind mbarriers): it is a
// correctness-first kernel whose point is to keep the N x N matrices out of memory -- the materialised per-head
// backward it replaces moved ~350 MB and 12 launches per head at 4096 tokens.
// Warp roles (160 threads): warps 0-3 = softmax / dQ / epilogue (warp w owns TMEM lanes 32 w ..), warp 4 = TMA + MMA.
#include "attention_bwd_sm100.cuh"

namespace unib {

namespace {
constexpr int kTile = 16384;                       // one [128 rows x 64 fp16] swizzled tile
constexpr int kColS = 0, kColDp = 128, kColDv = 256;          // dK at 256 + dpad, dQ at 256 + 2 dpad (3 dpad <= 256)
// NCH = 64-column chunks of the head dimension: 1 (d <= 64) or 2 (d <= 80: three accumulators of dpad columns must
// fit the 256 TMEM columns left beside S and dP)
template <int NCH>
struct BwdCfg {
  static constexpr int kOp = NCH * kTile;          // one [128 rows x d] operand tile
  static constexpr int kKOff = 0, kVOff = kOp, kQOff = 2 * kOp, kDoOff = 3 * kOp;
  static constexpr int kPOff = 4 * kOp, kDsOff = 4 * kOp + 2 * kTile;       // [128 q][128 k] = two 64-column chunks each
  static constexpr int kBarOff = 4 * kOp + 4 * kTile;
  static constexpr int kSmem = kBarOff + 128 + 1024;
  static_assert(kSmem <= 232448, "shared memory budget");
};
}  // namespace

template <int NCH>
__global__ void __launch_bounds__(160, 1)
attention_bwd_kernel(const __grid_constant__ AttnBwdMaps maps, const __grid_constant__ AttnBwdParams p) {
  using Cfg = BwdCfg<NCH>;
  constexpr int kKOff = Cfg::kKOff, kVOff = Cfg::kVOff, kQOff = Cfg::kQOff, kDoOff = Cfg::kDoOff;
  constexpr int kPOff = Cfg::kPOff, kDsOff = Cfg::kDsOff, kBarOff = Cfg::kBarOff;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar = base + kBarOff;
  const uint32_t bar_kv = bar, bar_q = bar + 8, bar_sdp = bar + 16, bar_p = bar + 24, bar_mma2 = bar + 32, bar_dq = bar + 40;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nq_tiles = (p.Nq + 127) / 128;
  const int dpad = (p.d + 15) & ~15;
  const int ks_d = dpad / 16;
  const int kColDk = kColDv + dpad, kColDq = kColDv + 2 * dpad;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    tma_prefetch_desc(&maps.dout);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_mma2, 1);
    mbar_init(bar_dq, 128);
    fence_barrier_init();
  }
  if (warp == 4) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // =============================== TMA + MMA (one elected lane) ===============================
    const uint32_t idesc_s = make_idesc_f16(128, 128);              // S, dP: A, B K-major
    const uint32_t idesc_t = make_idesc_f16(128, dpad, 1, 1);       // dV, dK: A = [q][k] tile transposed, B MN-major
    const uint32_t idesc_q = make_idesc_f16(128, dpad, 0, 1);       // dQ: A K-major, B = K block MN-major
    const uint64_t k_desc = make_desc_kmajor_sw128(base + kKOff), v_desc = make_desc_kmajor_sw128(base + kVOff);
    const uint64_t q_desc = make_desc_kmajor_sw128(base + kQOff), do_desc = make_desc_kmajor_sw128(base + kDoOff);
    const uint64_t pT_desc = make_desc_mnmajor_sw128(base + kPOff, kTile, 1024);
    const uint64_t dsT_desc = make_desc_mnmajor_sw128(base + kDsOff, kTile, 1024);
    const uint64_t ds_desc = make_desc_kmajor_sw128(base + kDsOff);
    const uint64_t doB_desc = make_desc_mnmajor_sw128(base + kDoOff, kTile, 1024);
    const uint64_t qB_desc = make_desc_mnmajor_sw128(base + kQOff, kTile, 1024);
    const uint64_t kB_desc = make_desc_mnmajor_sw128(base + kKOff, kTile, 1024);
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_kv, 2 * Cfg::kOp);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        tma_load_4d(base + kKOff + ch * kTile, &maps.k, bar_kv, ch * 64, jblk * 128, head, b);
        tma_load_4d(base + kVOff + ch * kTile, &maps.v, bar_kv, ch * 64, jblk * 128, head, b);
      }
    }
    mbar_wait(bar_kv, 0);
    for (int i = 0; i < nq_tiles; ++i) {
      const uint32_t ph = i & 1;
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, 2 * Cfg::kOp);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tma_load_4d(base + kQOff + ch * kTile, &maps.q, bar_q, ch * 64, i * 128, head, b);
          tma_load_4d(base + kDoOff + ch * kTile, &maps.dout, bar_q, ch * 64, i * 128, head, b);
        }
      }
      mbar_wait(bar_q, ph);
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < ks_d; ++ks) {                        // S = Q K^T, dP = dO V^T: K steps of 16 over d
          const uint64_t off = static_cast<uint64_t>(((ks >> 2) * kTile + (ks & 3) * 32) >> 4);
          umma_f16_ss(tmem_base + kColS, q_desc + off, k_desc + off, idesc_s, ks > 0 ? 1u : 0u);
          umma_f16_ss(tmem_base + kColDp, do_desc + off, v_desc + off, idesc_s, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_sdp);
      }
      mbar_wait(bar_p, ph);                                        // P, dS are in shared memory
      if (i > 0) mbar_wait(bar_dq, (i - 1) & 1);                   // dQ of the previous tile has been drained
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < 8; ++ks) {                           // contraction over the 128 query rows, 16 at a time
          const uint64_t step = static_cast<uint64_t>((ks * 2048) >> 4);
          umma_f16_ss(tmem_base + kColDv, pT_desc + step, doB_desc + step, idesc_t, (i > 0 || ks > 0) ? 1u : 0u);
          umma_f16_ss(tmem_base + kColDk, dsT_desc + step, qB_desc + step, idesc_t, (i > 0 || ks > 0) ? 1u : 0u);
        }
        for (int ks = 0; ks < 8; ++ks) {                           // dQ = dS K: contraction over the 128 keys
          const int ch = ks >> 2, within = ks & 3;
          umma_f16_ss(tmem_base + kColDq, ds_desc + static_cast<uint64_t>((ch * kTile + within * 32) >> 4),
                      kB_desc + static_cast<uint64_t>((ks * 2048) >> 4), idesc_q, ks > 0 ? 1u : 0u);
        }
        umma_commit(bar_mma2);
      }
      mbar_wait(bar_mma2, ph);                                     // Q / dO / P / dS buffers are free again
    }
  } else {
    // =============================== softmax, dQ, epilogue: thread = row ===============================
    const int row = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int kv_valid = p.Nk - jblk * 128;                        // columns >= kv_valid are padding
    const size_t bh = static_cast<size_t>(b) * p.heads + head;
    const int sw = row & 7;
    const uint32_t p_row = base + kPOff + row * 128, ds_row = base + kDsOff + row * 128;
    for (int i = 0; i < nq_tiles; ++i) {
      const uint32_t ph = i & 1;
      const int q = i * 128 + row;
      const bool q_ok = q < p.Nq;
      const float lse2 = q_ok ? p.lse2[bh * p.Nq + q] : 0.f;
      const float Dq = q_ok ? p.D[bh * p.Nq + q] : 0.f;
      mbar_wait(bar_sdp, ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {                                // 32 key columns at a time
        float s[32], dp[32];
        tmem_ld32(tmem_base + lane_off + kColS + c * 32, s);
        tmem_ld32(tmem_base + lane_off + kColDp + c * 32, dp);
        tmem_ld_wait();
        uint32_t pk[16], dk2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int col = c * 32 + 2 * e;
          float p0 = fast_exp2(s[2 * e] * sl2 - lse2), p1 = fast_exp2(s[2 * e + 1] * sl2 - lse2);
          if (!q_ok || col >= kv_valid) p0 = 0.f;
          if (!q_ok || col + 1 >= kv_valid) p1 = 0.f;
          pk[e] = pack_half2(p0, p1);
          dk2[e] = pack_half2(p0 * (dp[2 * e] - Dq) * p.scale, p1 * (dp[2 * e + 1] - Dq) * p.scale);
        }
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {                           // four 16-byte units of this 32-column slice
          const int u = c * 4 + u4;                                // unit index 0..15 over the 128 columns
          const uint32_t off = (u >> 3) * kTile + (((u & 7) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + off), "r"(pk[u4 * 4 + 0]),
                       "r"(pk[u4 * 4 + 1]), "r"(pk[u4 * 4 + 2]), "r"(pk[u4 * 4 + 3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + off), "r"(dk2[u4 * 4 + 0]),
                       "r"(dk2[u4 * 4 + 1]), "r"(dk2[u4 * 4 + 2]), "r"(dk2[u4 * 4 + 3]) : "memory");
        }
      }
      fence_proxy_async_shared();
      tc_fence_before();
      mbar_arrive(bar_p);
      // dQ_i of this key block -> fp32 accumulator
      mbar_wait(bar_mma2, ph);
      tc_fence_after();
      float* dq_row = p.dq_acc + (static_cast<size_t>(b) * p.Nq + q) * p.ld_dq + head * p.d;
#pragma unroll 1
      for (int c = 0; c < dpad / 16; ++c) {
        float o[16];
        tmem_ld16(tmem_base + lane_off + kColDq + c * 16, o);
        tmem_ld_wait();
        if (q_ok) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (c * 16 + e < p.d) atomicAdd(dq_row + c * 16 + e, o[e]);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_dq);
    }
    // dK, dV of this key block (complete: the last tile's bar_mma2 has been observed above)
    const int key = jblk * 128 + row;
    const bool key_ok = key < p.Nk;                                // TMEM loads are warp-collective: every lane loads
    __half* dk_row = p.dk + (static_cast<size_t>(b) * p.Nk + key) * p.ld_dk + head * p.d;
    __half* dv_row = p.dv + (static_cast<size_t>(b) * p.Nk + key) * p.ld_dv + head * p.d;
#pragma unroll 1
    for (int c = 0; c < dpad / 16; ++c) {
      float a[16], g[16];
      tmem_ld16(tmem_base + lane_off + kColDk + c * 16, a);
      tmem_ld16(tmem_base + lane_off + kColDv + c * 16, g);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int col = c * 16 + u * 8;
        if (key_ok && col + 8 <= p.d) {
          uint4 w1, w2;
          w1.x = pack_half2(a[u * 8 + 0], a[u * 8 + 1]); w1.y = pack_half2(a[u * 8 + 2], a[u * 8 + 3]);
          w1.z = pack_half2(a[u * 8 + 4], a[u * 8 + 5]); w1.w = pack_half2(a[u * 8 + 6], a[u * 8 + 7]);
          w2.x = pack_half2(g[u * 8 + 0], g[u * 8 + 1]); w2.y = pack_half2(g[u * 8 + 2], g[u * 8 + 3]);
          w2.z = pack_half2(g[u * 8 + 4], g[u * 8 + 5]); w2.w = pack_half2(g[u * 8 + 6], g[u * 8 + 7]);
          *reinterpret_cast<uint4*>(dk_row + col) = w1;
          *reinterpret_cast<uint4*>(dv_row + col) = w2;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[b, h, q] = sum_c dO[q, h*d + c] * O[q, h*d + c]  (fp32) -- one warp per (row, head)
__global__ void __launch_bounds__(256) attention_bwd_prep_kernel(const __half* __restrict__ o, int ldo,
                                                                 const __half* __restrict__ dout, int lddo, float* __restrict__ D,
                                                                 int B, int heads, int Nq, int d) {
  const long long total = static_cast<long long>(B) * heads * Nq;
  const int lane = threadIdx.x & 31;
  for (long long w = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; w < total;
       w += (static_cast<long long>(gridDim.x) * blockDim.x) >> 5) {
    const int q = static_cast<int>(w % Nq);
    const int h = static_cast<int>((w / Nq) % heads);
    const int b = static_cast<int>(w / (static_cast<long long>(Nq) * heads));
    const size_t r = static_cast<size_t>(b) * Nq + q;
    float acc = 0.f;
    for (int c = lane; c < d; c += 32)
      acc += __half2float(o[r * ldo + h * d + c]) * __half2float(dout[r * lddo + h * d + c]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) D[(static_cast<size_t>(b) * heads + h) * Nq + q] = acc;
  }
}

template <int NCH>
static cudaError_t launch_bwd_cfg(const AttnBwdMaps& maps, const AttnBwdParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         BwdCfg<NCH>::kSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.Nk + 127) / 128, p.heads, p.B);
  attention_bwd_kernel<NCH><<<grid, 160, BwdCfg<NCH>::kSmem, stream>>>(maps, p);
  return cudaGetLastError();
}

int attention_bwd_max_d() { return 80; }

cudaError_t launch_attention_bwd(const AttnBwdMaps& maps, const AttnBwdParams& p, const __half* o, int ldo,
                                 const __half* dout, int lddo, cudaStream_t stream) {
  if (p.d % 8 != 0 || p.d < 8 || p.d > attention_bwd_max_d()) return cudaErrorInvalidValue;
  const long long rows = static_cast<long long>(p.B) * p.heads * p.Nq;
  int blocks = static_cast<int>((rows * 32 + 255) / 256 > 148 * 16 ? 148 * 16 : (rows * 32 + 255) / 256);
  attention_bwd_prep_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(o, ldo, dout, lddo, p.D, p.B, p.heads, p.Nq, p.d);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return p.d <= 64 ? launch_bwd_cfg<1>(maps, p, stream) : launch_bwd_cfg<2>(maps, p, stream);
}

}  // namespace unib
