// Implicit-GEMM convolution / linear kernel for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (cta_group::1,
// UMMA 128 x BN x 16, fp16 in / fp32 accumulate in TMEM, double-buffered accumulators) -> tcgen05.ld epilogue.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue.
// Persistent: each CTA walks work items (split, m_tile, n_tile) with stride gridDim.x.
//
// Epilogue (NHWC fp16 outputs): the 128 x BN tile is drained in 32-column sub-tiles through a ring of 64B-swizzled
// shared-memory slots.  The residual sub-tile (ResNet identity / transformer skip / exchange add) is TMA-LOADED into
// the slot a few sub-tiles ahead, each thread adds its own row in place (bias from a per-tile smem copy, SiLU / GEGLU
// gate in registers) and the slot is TMA-STORED to HBM -- coalesced 64 B rows, no per-thread global latency on the
// critical path.  fp32 split-K partials and the tiny NCHW outputs (conv_out + fused scheduler update) use direct
// per-thread stores.
#include "gemm_sm100.cuh"

namespace unib {

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSlots = (BN >= 256) ? 3 : 4;           // epilogue slot ring (8 KB each)
  static constexpr int kLookahead = kSlots - 2;                  // residual sub-tiles requested ahead
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 160) ? 5 : 6;
  static constexpr int kSlotBytes = kBM * 64;                    // [128 rows x 32 fp16], SWIZZLE_64B
  static constexpr int kEpiOff = kStages * kStageBytes;
  static constexpr int kBiasOff = kEpiOff + kSlots * kSlotBytes; // [2][BN] fp32
  static constexpr int kBarOff = kBiasOff + 2 * BN * 4;
  static constexpr int kNumBars = 2 * kStages + 4 + kSlots;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kBarOff + kNumBars * 8 + 16 + 1024;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N constraint for M=128 / 32-column epilogue sub-tiles");
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment for SWIZZLE_128B");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct WorkItem {
  int mt, nt, kb0, kb1;
};
__device__ __forceinline__ WorkItem decode_work(const GemmParams& p, int w) {
  WorkItem wi;
  const int tiles = p.m_tiles * p.n_tiles;
  const int split = w / tiles;
  const int rem = w - split * tiles;
  wi.mt = rem / p.n_tiles;
  wi.nt = rem - wi.mt * p.n_tiles;
  wi.kb0 = static_cast<int>((static_cast<long long>(split) * p.total_kb) / p.splits);
  wi.kb1 = static_cast<int>((static_cast<long long>(split + 1) * p.total_kb) / p.splits);
  return wi;
}

// 16 accumulator columns [n, n+16) of row m -> final output (direct-store path: NCHW outputs, split-K finalize,
// shapes the TMA epilogue does not cover).  `v` already holds fp32 sums.
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, float* v, int m, int n) {
  const int N = p.N;
  if (n >= N) return;
  const int b = m / p.rows_per_batch;
  if (p.bias != nullptr) {
    const float* bp = p.bias + static_cast<size_t>(b) * p.bias_bstride + n;
    if (n + 16 <= N) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(bp + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    } else {
      for (int j = 0; j < 16 && n + j < N; ++j) v[j] += bp[j];
    }
  }
  if (p.flags & EPI_OUT_NCHW) {
    const int hw = m - b * p.rows_per_batch;
    for (int j = 0; j < 16 && n + j < N; ++j) {
      const size_t idx = (static_cast<size_t>(b) * N + (n + j)) * p.rows_per_batch + hw;
      float val = v[j];
      if (p.flags & EPI_AXPBY) {
        const float* cf = p.axpby + (p.axpby_step ? 2 * static_cast<size_t>(*p.axpby_step) : 0);
        const float x = p.aux[idx];
        const bool keep = n + j < p.axpby_n0;
        // optional NHWC fp16 copy of the raw prediction with the clean channels passed through: the attribute
        // input of a follow-up pass (cat(latents_mask, mask_pred), train/train.py:1393)
        if (p.out != nullptr)
          reinterpret_cast<__half*>(p.out)[static_cast<size_t>(m) * p.ldc + n + j] = __float2half_rn(keep ? x : val);
        p.aux_out[idx] = keep ? x : cf[0] * val + cf[1] * x;
      } else if (p.flags & EPI_OUT_F32) {
        reinterpret_cast<float*>(p.out)[idx] = val;
      } else {
        reinterpret_cast<__half*>(p.out)[idx] = __float2half_rn(val);
      }
    }
    return;
  }
  __half* op = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(m) * p.ldc + n;
  if (n + 16 <= N) {
    if (p.res != nullptr) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.res + static_cast<size_t>(m) * p.ldr + n);
      const uint4 r0 = rp[0], r1 = rp[1];
      const __half2* h0 = reinterpret_cast<const __half2*>(&r0);
      const __half2* h1 = reinterpret_cast<const __half2*>(&r1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f0 = __half22float2(h0[j]), f1 = __half22float2(h1[j]);
        v[2 * j] += f0.x; v[2 * j + 1] += f0.y;
        v[8 + 2 * j] += f1.x; v[8 + 2 * j + 1] += f1.y;
      }
    }
    if (p.flags & EPI_SILU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
    }
    uint4 o0, o1;
    o0.x = pack_half2(v[0], v[1]);   o0.y = pack_half2(v[2], v[3]);
    o0.z = pack_half2(v[4], v[5]);   o0.w = pack_half2(v[6], v[7]);
    o1.x = pack_half2(v[8], v[9]);   o1.y = pack_half2(v[10], v[11]);
    o1.z = pack_half2(v[12], v[13]); o1.w = pack_half2(v[14], v[15]);
    reinterpret_cast<uint4*>(op)[0] = o0;
    reinterpret_cast<uint4*>(op)[1] = o1;
  } else {
    for (int j = 0; j < 16 && n + j < N; ++j) {
      float val = v[j];
      if (p.res != nullptr) val += __half2float(p.res[static_cast<size_t>(m) * p.ldr + n + j]);
      if (p.flags & EPI_SILU) val = silu_f(val);
      op[j] = __float2half_rn(val);
    }
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + Cfg::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  auto res_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 4 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kBarOff + Cfg::kNumBars * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    if (p.epi_tma) {
      tma_prefetch_desc(&maps.c);
      if (p.res != nullptr) tma_prefetch_desc(&maps.r);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    for (int s = 0; s < Cfg::kSlots; ++s) mbar_init(res_bar(s), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_work = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkItem wi = decode_work(p, w);
        const int p0 = wi.mt * kBM;
        const int w0 = p0 % p.W;
        const int h0 = (p0 / p.W) % p.H;
        const int b0 = p0 / (p.W * p.H);
        int seg = 0, tap = 0, cb = 0;
        {
          int k = wi.kb0;
          while (seg < p.nseg) {
            const int per = p.seg[seg].ntaps * p.seg[seg].nkb;
            if (k < per) { tap = k / p.seg[seg].nkb; cb = k - tap * p.seg[seg].nkb; break; }
            k -= per;
            ++seg;
          }
        }
        for (int kb = wi.kb0; kb < wi.kb1; ++kb, ++it) {
          const int stage = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(empty_bar(stage), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
          const ConvSeg sg = p.seg[seg];
          int dw = 0, dh = 0, tm = sg.tmap;
          if (sg.kind == SEG_3x3) {
            dh = tap / 3 - 1;
            dw = tap % 3 - 1;
          } else if (sg.kind == SEG_3x3_S2) {
            const int dy = tap / 3, dx = tap % 3;   // input row = 2*oh + dy - 1 -> parity (dy != 1), half-res shift
            tm += ((dy != 1) ? 2 : 0) + ((dx != 1) ? 1 : 0);
            dh = (dy == 0) ? -1 : 0;
            dw = (dx == 0) ? -1 : 0;
          }
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          tma_load_4d(a_dst, &maps.a[tm], full_bar(stage), cb * kBK, w0 + dw, h0 + dh, b0);
          tma_load_2d(a_dst + Cfg::kABytes, &maps.b, full_bar(stage), kb * kBK, wi.nt * BN);
          if (++cb == sg.nkb) {
            cb = 0;
            if (++tap == sg.ntaps) { tap = 0; ++seg; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBM, BN);
      uint32_t it = 0, tl = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tl) {
        const WorkItem wi = decode_work(p, w);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = wi.kb0; kb < wi.kb1; ++kb, ++it) {
          const int stage = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(full_bar(stage), ph);
          tc_fence_after();
          const uint32_t a_addr = base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = make_desc_kmajor_sw128(a_addr);
          const uint64_t b_desc = make_desc_kmajor_sw128(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 B per UMMA_K=16 step inside the 128 B swizzle row (descriptor address is in 16 B units)
            umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > wi.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs have read it
        }
        umma_commit(tfull_bar(acc));        // accumulator complete -> epilogue
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;        // 0..127 within the epilogue group
    const bool leader = (et == 0);
    uint32_t tl = 0;
    if (p.epi_tma) {
      // ---------------- TMA epilogue: slot ring, residual prefetch, bias in smem ----------------
      const bool geglu = (p.flags & EPI_GEGLU) != 0;
      const bool silu = (p.flags & EPI_SILU) != 0;
      const bool has_res = p.res != nullptr;
      const int nsub = geglu ? BN / 64 : BN / 32;             // output sub-tiles per tile
      const int out_bn = geglu ? BN / 2 : BN;                 // output columns per tile
      float* bias_s = reinterpret_cast<float*>(smem + Cfg::kBiasOff);
      const uint32_t slot0 = base + Cfg::kEpiOff;
      const uint32_t my_row_off = static_cast<uint32_t>(row) * 64u;
      const uint32_t sw = static_cast<uint32_t>((row >> 1) & 3);
      uint32_t gsub = 0;                                       // sub-tiles drained so far (slot = gsub % kSlots)
      // residual prefetch cursor (leader only): runs kLookahead sub-tiles ahead of gsub
      int pf_w = blockIdx.x, pf_j = 0;
      uint32_t pf_g = 0;
      auto prefetch_res = [&]() {
        if (pf_w >= total_work) return;
        const WorkItem pw = decode_work(p, pf_w);
        const int s = pf_g % Cfg::kSlots;
        mbar_arrive_expect_tx(res_bar(s), Cfg::kSlotBytes);
        tma_load_2d(slot0 + s * Cfg::kSlotBytes, &maps.r, res_bar(s), pw.nt * out_bn + pf_j * 32, pw.mt * kBM);
        ++pf_g;
        if (++pf_j == nsub) { pf_j = 0; pf_w += gridDim.x; }
      };
      if (leader && has_res) {
        for (int i = 0; i < Cfg::kLookahead; ++i) prefetch_res();
      }
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tl) {
        const WorkItem wi = decode_work(p, w);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        const int m0 = wi.mt * kBM;
        const int m = m0 + row;
        // bias of this tile -> smem (two batch rows: a tile may straddle a batch boundary); issued before the
        // accumulator wait so the global latency overlaps the mainloop
        const bool per_batch = p.bias_bstride != 0;
        const int b_first = per_batch ? m0 / p.rows_per_batch : 0;
        int m_last = m0 + kBM - 1;
        if (m_last >= p.M) m_last = p.M - 1;
        const int b_last = per_batch ? m_last / p.rows_per_batch : 0;
        const bool bias_smem = p.bias != nullptr && (b_last - b_first) <= 1;
        float bv[2][(BN + 127) / 128];
        if (bias_smem) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < (BN + 127) / 128; ++c) {
              const int col = c * 128 + et;
              const int n = wi.nt * BN + col;
              const int bb = r == 0 ? b_first : b_last;
              bv[r][c] = (col < BN && n < p.N) ? __ldg(p.bias + static_cast<size_t>(bb) * p.bias_bstride + n) : 0.f;
            }
        }
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        if (bias_smem) {
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < (BN + 127) / 128; ++c) {
              const int col = c * 128 + et;
              if (col < BN) bias_s[r * BN + col] = bv[r][c];
            }
        }
        epi_bar_sync();     // bias visible; also orders the previous tile's last slot reads before new writes
        const int my_b = per_batch ? m / p.rows_per_batch : 0;
        const float* my_bias = bias_s + ((my_b > b_first) ? BN : 0);
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
        for (int j = 0; j < nsub; ++j, ++gsub) {
          const int s = gsub % Cfg::kSlots;
          const uint32_t slot = slot0 + s * Cfg::kSlotBytes + my_row_off;
          float v[32];
          if (geglu) {
            float gte[32];
            tmem_ld32(taddr + j * 32, v);
            tmem_ld32(taddr + BN / 2 + j * 32, gte);
            tmem_ld_wait();
            if (j == nsub - 1) {              // accumulator fully read -> release it to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty_bar(acc));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float a = v[i], g = gte[i];
              if (bias_smem) { a += my_bias[j * 32 + i]; g += my_bias[BN / 2 + j * 32 + i]; }
              v[i] = a * gelu_erf_f(g);
            }
          } else {
            tmem_ld32(taddr + j * 32, v);
            tmem_ld_wait();
            if (j == nsub - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty_bar(acc));
            }
            if (bias_smem) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += my_bias[j * 32 + i];
            } else if (p.bias != nullptr && m < p.M) {
              const float* bp = p.bias + static_cast<size_t>(my_b) * p.bias_bstride + wi.nt * BN + j * 32;
              for (int i = 0; i < 32; ++i)
                if (wi.nt * BN + j * 32 + i < p.N) v[i] += __ldg(bp + i);
            }
          }
          if (has_res) {
            mbar_wait(res_bar(s), (gsub / Cfg::kSlots) & 1);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t r0, r1, r2, r3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                           : "r"(slot + ((static_cast<uint32_t>(u) ^ sw) << 4)));
              const uint32_t rr[4] = {r0, r1, r2, r3};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rr[e]));
                v[u * 8 + 2 * e] += f.x;
                v[u * 8 + 2 * e + 1] += f.y;
              }
            }
          }
          if (silu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + ((static_cast<uint32_t>(u) ^ sw) << 4)),
                         "r"(pack_half2(v[u * 8 + 0], v[u * 8 + 1])), "r"(pack_half2(v[u * 8 + 2], v[u * 8 + 3])),
                         "r"(pack_half2(v[u * 8 + 4], v[u * 8 + 5])), "r"(pack_half2(v[u * 8 + 6], v[u * 8 + 7]))
                         : "memory");
          }
          fence_proxy_async_shared();
          epi_bar_sync();
          if (leader) {
            tma_store_2d(slot0 + s * Cfg::kSlotBytes, &maps.c, wi.nt * out_bn + j * 32, m0);
            tma_store_commit();
            tma_store_wait_read<1>();       // every store but the newest has finished reading its slot
            if (has_res) prefetch_res();
          }
        }
      }
      if (leader) tma_store_wait_all();
    } else {
      // ---------------- direct-store epilogue: split-K partials, NCHW outputs ----------------
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tl) {
        const WorkItem wi = decode_work(p, w);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        const int m = wi.mt * kBM + row;
        const bool row_ok = m < p.M;
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        if (p.splits > 1) {
          const int split = w / (p.m_tiles * p.n_tiles);
          float* pp = p.partial + (static_cast<size_t>(split) * p.M + m) * p.N + wi.nt * BN;
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            const int n = wi.nt * BN + c * 32;
            if (row_ok) {
              if (n + 32 <= p.N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(pp + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
                for (int j = 0; j < 32 && n + j < p.N; ++j) pp[c * 32 + j] = v[j];
              }
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < BN / 16; ++c) {
            float v[16];
            tmem_ld16(taddr + c * 16, v);
            tmem_ld_wait();
            if (row_ok) epilogue_store16(p, v, m, wi.nt * BN + c * 16);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Split-K finalize: sum fp32 partials over splits in fixed order, then the same epilogue as the fused path.
__global__ void __launch_bounds__(256) gemm_splitk_finalize_kernel(const GemmParams p) {
  const int chunks = (p.N + 15) / 16;
  const long long total = static_cast<long long>(p.M) * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / chunks);
    const int n = static_cast<int>(i - static_cast<long long>(m) * chunks) * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    for (int s = 0; s < p.splits; ++s) {
      const float* pp = p.partial + (static_cast<size_t>(s) * p.M + m) * p.N + n;
      if (n + 16 <= p.N) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(pp + j);
          v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
        }
      } else {
        for (int j = 0; j < 16 && n + j < p.N; ++j) v[j] += pp[j];
      }
    }
    epilogue_store16(p, v, m, n);
  }
}

size_t gemm_smem_bytes(int bn) {
  switch (bn) {
    case 32: return GemmCfg<32>::kSmemBytes;
    case 64: return GemmCfg<64>::kSmemBytes;
    case 128: return GemmCfg<128>::kSmemBytes;
    case 160: return GemmCfg<160>::kSmemBytes;
    case 256: return GemmCfg<256>::kSmemBytes;
  }
  return 0;
}

int gemm_pick_bn(int N, int flags) {
  if (flags & EPI_GEGLU) {                       // value/gate halves must be whole 32-column sub-tiles
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i)
      if (N % cands[i] == 0) return cands[i];
    return 0;
  }
  const int cands[4] = {160, 128, 64, 32};
  for (int i = 0; i < 4; ++i)
    if (N % cands[i] == 0) return cands[i];
  return N >= 128 ? 128 : (N > 32 ? 64 : 32);
}

template <int BN>
static cudaError_t launch_bn(const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GemmCfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int total_work = p.m_tiles * p.n_tiles * p.splits;
  const int grid = total_work < num_sms ? total_work : num_sms;
  gemm_tcgen05_kernel<BN><<<grid, 192, GemmCfg<BN>::kSmemBytes, stream>>>(maps, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (p.splits > 1) {
    const long long total = static_cast<long long>(p.M) * ((p.N + 15) / 16);
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    gemm_splitk_finalize_kernel<<<blocks, 256, 0, stream>>>(p);
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_gemm(const GemmMaps& maps, const GemmParams& p, int bn, int num_sms, cudaStream_t stream) {
  switch (bn) {
    case 32: return launch_bn<32>(maps, p, num_sms, stream);
    case 64: return launch_bn<64>(maps, p, num_sms, stream);
    case 128: return launch_bn<128>(maps, p, num_sms, stream);
    case 160: return launch_bn<160>(maps, p, num_sms, stream);
    case 256: return launch_bn<256>(maps, p, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace unib
