// Implicit-GEMM convolution / linear kernel for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (UMMA
// 128 x BN x 16 per CTA, or 256 x BN x 16 per CTA PAIR with cta_group::2; fp16 in / fp32 accumulate in TMEM,
// double-buffered accumulators) -> tcgen05.ld epilogue.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 = epilogue.
// Persistent: each CTA (pair) walks work items (split, m_tile, n_tile) with stride gridDim.x (/ 2).
//
// Epilogue (NHWC fp16 outputs): the eight epilogue warps (two per TMEM lane quadrant = 32 rows of the 128 x BN tile,
// splitting its 32-column sub-tiles) drain the accumulator straight from TMEM through registers: bias from a per-warp smem copy, the residual (ResNet identity /
// transformer skip / exchange add) read with 256-bit global loads one sub-tile ahead, SiLU / GEGLU gate in registers,
// 256-bit global stores (64 contiguous bytes per row and sub-tile = full sectors).  No smem staging, proxy fence or
// block-wide barrier on the critical path (a TMA-store slot ring measured ~1000 cycles of serial latency per
// sub-tile and made every small-K GEMM epilogue-bound).  fp32 split-K partials and the tiny NCHW outputs (conv_out +
// fused scheduler update) use direct per-thread stores.
#include "gemm_sm100.cuh"

#include <cstdlib>
#include <type_traits>

namespace unib {

// CG = CTAs per MMA: 1 = cta_group::1 (tile 128 x BN per CTA); 2 = cta_group::2 (tile 256 x BN per CTA PAIR: each CTA
// stages its own 128 rows of A and HALF of the B tile, so the per-SM operand ingress per K block drops from
// 16 KB + BN*128 B to 16 KB + BN*64 B -- the measured limiter of the mainloop)
template <int BN, int CG = 1>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // per epilogue warp (8): bias [2 batch rows] + LayerNorm wsum [1] for ITS half of the tile's 32-column sub-tiles, fp32
  // (GEGLU keeps value and gate columns: 2 x the slots of its BN / 64 output sub-tiles)
  static constexpr int kVecFloats = 3 * (((BN / 32 + 1) / 2) * 32);
  static constexpr int kGegluFloats = 6 * (((BN / 64 + 1) / 2) * 32);
  static constexpr int kBiasBytes = 8 * 4 * (kVecFloats > kGegluFloats ? kVecFloats : kGegluFloats);
  // GroupNorm statistics fused into the epilogue: pair partials of one tile, [2 tile parities][4 lane quadrants][BN]
  static constexpr int kStatBytes = 2 * 4 * BN * 4;
  static constexpr int kStagesFit = (232448 - 1024 - 512 - kBiasBytes - kStatBytes) / kStageBytes;
#ifndef UNIB_MAX_STAGES
#define UNIB_MAX_STAGES 8
#endif
  static constexpr int kStages = kStagesFit > UNIB_MAX_STAGES ? UNIB_MAX_STAGES : kStagesFit;
  static constexpr int kBiasOff = kStages * kStageBytes;
  static constexpr int kStatOff = kBiasOff + kBiasBytes;
  static constexpr int kBarOff = kStatOff + kStatBytes;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kBarOff + kNumBars * 8 + 16 + 1024;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "UMMA N constraint for M=128 / 32-column epilogue sub-tiles");
  static_assert(CG == 1 || CG == 2, "one CTA or a CTA pair");
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment for SWIZZLE_128B");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct WorkItem {
  int mt, nt, kb0, kb1;
};
// Integer division has no hardware unit: a runtime-divisor `/` costs ~100+ cycles, and the work decode of every tile
// (three roles) plus the tile-origin decode used to burn ~0.9 us at the head of every launch.  Divisors are fixed per
// launch, so the host precomputes multiply-shift reciprocals (exact for dividends and divisors < 2^20) and shifts for
// the power-of-two image dims.
__device__ __forceinline__ int fast_div(int n, unsigned long long mul) {       // n / d with mul = floor(2^40 / d) + 1
  return static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(n)) * mul) >> 40);
}
__device__ __forceinline__ int batch_of_row(const GemmParams& p, int m) {
  return p.rpb_shift >= 0 ? (m >> p.rpb_shift) : m / p.rows_per_batch;
}
__device__ __forceinline__ WorkItem decode_work(const GemmParams& p, int w) {
  WorkItem wi;
  const int tiles = p.m_tiles * p.n_tiles;
  int split = 0, rem = w;
  if (p.splits > 1) {
    split = fast_div(w, p.mul_tiles);
    rem = w - split * tiles;
  }
  wi.mt = fast_div(rem, p.mul_ntiles);
  wi.nt = rem - wi.mt * p.n_tiles;
  if (p.splits > 1) {
    wi.kb0 = static_cast<int>((static_cast<long long>(split) * p.total_kb) / p.splits);
    wi.kb1 = static_cast<int>((static_cast<long long>(split + 1) * p.total_kb) / p.splits);
  } else {
    wi.kb0 = 0;
    wi.kb1 = p.total_kb;
  }
  return wi;
}

// 16 accumulator columns [n, n+16) of row m -> final output (direct-store path: NCHW outputs, split-K finalize,
// shapes the TMA epilogue does not cover).  `v` already holds fp32 sums.
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, float* v, int m, int n) {
  const int N = p.N;
  if (n >= N) return;
  const int b = batch_of_row(p, m);
  if (p.bias != nullptr) {
    const float* bp = p.bias + (p.bias_step ? static_cast<size_t>(*p.bias_step) * p.bias_step_stride : 0) +
                      static_cast<size_t>(b) * p.bias_bstride + n;
    if (n + 16 <= N) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(bp + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n + j < N) v[j] += bp[j];
    }
  }
  if (p.flags & EPI_OUT_NCHW) {
    const int hw = m - b * p.rows_per_batch;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (n + j >= N) break;
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
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (n + j >= N) break;
      float val = v[j];
      if (p.res != nullptr) val += __half2float(p.res[static_cast<size_t>(m) * p.ldr + n + j]);
      if (p.flags & EPI_SILU) val = silu_f(val);
      op[j] = __float2half_rn(val);
    }
  }
}

// debug stamps of CTA 0: trace[16 + slot * 16 + k] = globaltimer (ns); slot = launch ordinal.  Compiled in only with
// -DUNIB_GEMM_TRACE (tools/gemm_trace.py builds that variant): the producer and MMA-issue loops are single-thread
// latency chains, and even a predicated-off stamp inside them costs measurable mainloop throughput.
#ifdef UNIB_GEMM_TRACE
#define GEMM_TRACE(k)                                                                      \
  do {                                                                                     \
    if (p.trace != nullptr && blockIdx.x == 0 && *trace_slot < 4000)                       \
      p.trace[16 + *trace_slot * 16 + (k)] = static_cast<long long>(global_timer_ns());    \
  } while (0)
#else
#define GEMM_TRACE(k) do { } while (0)
#endif

// Epilogue modes (compile-time, so that each kernel's hot epilogue loop is short straight-line code: the one-kernel
// version with run-time flags spent 34 % of its issue slots on instruction-cache misses, ncu stall_no_inst):
//   MODE_VEC     NHWC fp16 output through registers -> 256-bit global stores; N a multiple of 32, rows 32 B aligned;
//                bias / residual / LayerNorm fold / row statistics are short run-time-optional blocks
//   MODE_GEGLU   value * gelu(gate) of the two halves of the tile, otherwise as MODE_VEC
//   MODE_DIRECT  everything else through per-thread scalar / 128-bit stores: split-K fp32 partials, NCHW outputs and
//                the fused scheduler update, SiLU, odd N, unaligned rows
enum GemmMode : int { MODE_VEC = 0, MODE_GEGLU = 1, MODE_DIRECT = 2 };

constexpr int kEpiWarps = 8;                       // two per TMEM lane quadrant: they split the tile's column sub-tiles
constexpr int kGemmThreads = (2 + kEpiWarps) * 32;

template <int BN, int CG, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, CG>;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;     // 0 = leader of the CTA pair (issues the MMAs)
  const int cta = blockIdx.x / CG, ncta = gridDim.x / CG;      // persistent walk: both CTAs of a pair share `cta`
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + Cfg::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kBarOff + Cfg::kNumBars * 8);
  volatile int* trace_slot = reinterpret_cast<volatile int*>(smem + Cfg::kBarOff + Cfg::kNumBars * 8 + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

#ifdef UNIB_GEMM_TRACE
  if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    *trace_slot = static_cast<int>(atomicAdd(reinterpret_cast<unsigned long long*>(p.trace), 1ull));
    GEMM_TRACE(0);
  }
#else
  (void)trace_slot;
#endif
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    if (p.dual) tma_prefetch_desc(&maps.b2);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps * CG);  // pair: the epilogue warps of BOTH CTAs release the leader's MMA warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_2cta(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();             // peer barriers initialised + TMEM allocated in both CTAs
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) GEMM_TRACE(1);
  // PDL: wait for the previous kernel (producer of our activations / residual / bias, last reader of our output
  // buffer) to complete; everything above overlapped its tail
  if (p.pdl_early) pdl_launch();
  pdl_wait();
  if (threadIdx.x == 0) GEMM_TRACE(14);

  const int total_work = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The WHOLE warp walks the loop (warp-uniform control flow keeps stage / coordinates in uniform registers, so
    // UTMALDG takes them directly); one elected lane issues.  The loop body is a single-thread latency chain that
    // paces the whole mainloop, so everything that only changes per tap / per tile is hoisted out of it.
    uint32_t stage = 0, ph = 0;
    for (int w = cta; w < total_work; w += ncta) {
      const WorkItem wi = decode_work(p, w);
      const int grp = (p.dual && wi.mt >= p.mt_single) ? 1 : 0;       // dual launch: which of the two problems
      const int p0 = ((wi.mt - grp * p.mt_single) * CG + static_cast<int>(rank)) * kBM;
      const int w0 = p0 & ((1 << p.w_shift) - 1);                     // W, H are powers of two (linear: W = 2^30)
      const int h0 = (p0 >> p.w_shift) & ((1 << p.h_shift) - 1);
      const int b0 = p0 >> (p.w_shift + p.h_shift);
      int seg = 0, tap = 0, cb = 0;
      if (wi.kb0 > 0) {                                               // split-K only: locate the first K block
        int k = wi.kb0;
        while (seg < p.nseg) {
          const int per = p.seg[seg].ntaps * p.seg[seg].nkb;
          if (k < per) { tap = k / p.seg[seg].nkb; cb = k - tap * p.seg[seg].nkb; break; }
          k -= per;
          ++seg;
        }
      }
      int nkb = 1, ntaps = 1, cw = w0, ch = h0;
      const CUtensorMap* amap = &maps.a[0];
      const int par = p.up_shift >= 0 ? (wi.nt >> p.up_shift) : 0;    // SEG_UP2x2: output parity (py, px) of this N tile
      auto set_tap = [&]() {                   // geometry of (seg, tap): tensor map + shifted tile origin
        const ConvSeg sg = p.seg[seg < p.nseg ? seg : 0];
        nkb = sg.nkb;
        ntaps = sg.ntaps;
        int dw = 0, dh = 0, tm = sg.tmap;
        if (sg.kind == SEG_3x3) {
          dh = tap / 3 - 1;
          dw = tap % 3 - 1;
        } else if (sg.kind == SEG_3x3_S2) {
          const int dy = tap / 3, dx = tap % 3;   // input row = 2*oh + dy - 1 -> parity (dy != 1), half-res shift
          tm += ((dy != 1) ? 2 : 0) + ((dx != 1) ? 1 : 0);
          dh = (dy == 0) ? -1 : 0;
          dw = (dx == 0) ? -1 : 0;
        } else if (sg.kind == SEG_3x3_S2P0) {
          const int dy = tap / 3, dx = tap % 3;   // input row = 2*oh + dy (pad bottom/right only) -> parity dy & 1
          tm += ((dy & 1) ? 2 : 0) + ((dx & 1) ? 1 : 0);
          dh = dy >> 1;
          dw = dx >> 1;
        } else if (sg.kind == SEG_UP2x2) {
          // nearest-2x + conv3x3 for output pixel (2h + py, 2w + px): the 3 x 3 taps on the upsampled image touch only
          // the 2 x 2 low-resolution pixels (h - 1 + py + ty, w - 1 + px + tx); their weights are pre-summed on the host
          dh = (tap >> 1) - 1 + (par >> 1);
          dw = (tap & 1) - 1 + (par & 1);
        }
        amap = &maps.a[tm + grp];
        cw = w0 + dw;
        ch = h0 + dh;
      };
      const CUtensorMap* bmap = grp ? &maps.b2 : &maps.b;
      set_tap();
      const int brow = wi.nt * BN + static_cast<int>(rank) * (BN / CG);   // pair: this CTA's half of the B tile
#ifdef UNIB_GEMM_TRACE
      if (w == cta && lane == 0) GEMM_TRACE(15);
#endif
      for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
        mbar_wait(empty_bar(stage), ph ^ 1);
        if (elect_one()) {
#ifdef UNIB_GEMM_TRACE
          if (w == cta && kb == wi.kb0) GEMM_TRACE(2);
#endif
          const uint32_t a_dst = base + stage * Cfg::kStageBytes;
          if (CG == 2) {
            // both CTAs' loads complete on the LEADER's full barrier, which expects the bytes of the whole pair
            const uint32_t full_leader = mapa_u32(full_bar(stage), 0);
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            tma_load_4d_2cta(a_dst, amap, full_leader, cb * kBK, cw, ch, b0);
            tma_load_2d_2cta(a_dst + Cfg::kABytes, bmap, full_leader, kb * kBK, brow);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_4d(a_dst, amap, full_bar(stage), cb * kBK, cw, ch, b0);
            tma_load_2d(a_dst + Cfg::kABytes, bmap, full_bar(stage), kb * kBK, brow);
          }
        }
        if (++cb == nkb) {
          cb = 0;
          if (++tap == ntaps) { tap = 0; ++seg; }
          set_tap();
        }
        if (++stage == Cfg::kStages) { stage = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // Same structure: warp-uniform loop, one elected lane issues the four UMMAs of a K block and the commit.
    constexpr uint32_t idesc = make_idesc_f16(kBM * CG, BN);
    uint32_t stage = 0, ph = 0, tl = 0;
    if (rank == 0) {                             // pair: only the leader CTA issues (one UMMA of M = 256 drives both)
      for (int w = cta; w < total_work; w += ncta, ++tl) {
        const WorkItem wi = decode_work(p, w);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
          mbar_wait(full_bar(stage), ph);
          tc_fence_after();
          if (elect_one()) {
#ifdef UNIB_GEMM_TRACE
            if (w == cta && kb == wi.kb0) GEMM_TRACE(3);
#endif
            const uint32_t a_addr = base + stage * Cfg::kStageBytes;
            const uint64_t a_desc = make_desc_kmajor_sw128(a_addr);
            const uint64_t b_desc = make_desc_kmajor_sw128(a_addr + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // +32 B per UMMA_K=16 step inside the 128 B swizzle row (descriptor address is in 16 B units)
              if (CG == 2) umma_f16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > wi.kb0 || k > 0) ? 1u : 0u);
              else umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > wi.kb0 || k > 0) ? 1u : 0u);
            }
            if (CG == 2) {
              umma_commit_2cta(empty_bar(stage), 3);                        // frees the stage in BOTH CTAs
              if (kb == wi.kb1 - 1) umma_commit_2cta(tfull_bar(acc), 3);    // both CTAs' epilogues
            } else {
              umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs have read it
              if (kb == wi.kb1 - 1) umma_commit(tfull_bar(acc));   // accumulator complete -> epilogue
            }
          }
          if (++stage == Cfg::kStages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue (warps 2..9) ===============================
    // Eight warps: warp w may access TMEM lane quadrant w & 3 (hardware rule), so two warps share each quadrant = 32
    // rows of the tile and split its 32-column sub-tiles between them (even / odd, flipped every tile so that odd
    // sub-tile counts balance).  Two epilogue warps per scheduler also hide each other's TMEM / global latencies.
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int hw = (warp - 2) >> 2;         // which half of the sub-tiles (before the per-tile flip)
    const int ew = warp - 2;                // 0..7
    // hand the accumulator back to the MMA warp -- which, for a CTA pair, lives in the leader CTA
    auto release_acc = [&](int a) {
      if (CG == 2 && rank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar(a), 0));
      else mbar_arrive(tempty_bar(a));
    };
    const bool leader = (threadIdx.x == 64);
    uint32_t tl = 0;
    if (MODE == MODE_VEC || MODE == MODE_GEGLU) {
      // ---------------- NHWC fp16 epilogue: registers -> 256-bit global stores ----------------
      // A thread owns one row of the tile and drains its warp's sub-tiles: tcgen05.ld (32 fp32 columns) -> LayerNorm
      // fold -> + bias (per-warp smem copy) -> + residual (two 256-bit global loads issued one sub-tile ahead) ->
      // [GEGLU gate] -> two 256-bit global stores (64 contiguous bytes = 2 full sectors per row).  No shared-memory
      // staging, no proxy fence, no block-wide barrier: the warps run fully decoupled.
      constexpr bool geglu = (MODE == MODE_GEGLU);
      constexpr int nsub = geglu ? BN / 64 : BN / 32;           // output sub-tiles per tile
      constexpr int nmine = (nsub + 1) / 2;                     // at most this many per warp
      constexpr int out_bn = geglu ? BN / 2 : BN;               // output columns per tile
      constexpr int kWarpFloats = (geglu ? 3 * 2 : 3) * nmine * 32;
      static_assert(kEpiWarps * kWarpFloats * 4 <= Cfg::kBiasBytes, "per-warp bias copies must fit");
#if defined(UNIB_WHATIF_EPI) && (UNIB_WHATIF_EPI & 2)
      const bool has_res = false;                // what-if timing aid: the epilogue without its residual loads
#else
      const bool has_res = !geglu && p.res != nullptr;
#endif
      const bool has_bias = p.bias != nullptr;
      const int n_out = geglu ? p.N / 2 : p.N;
      // this warp's bias copy: [2 batch rows][slots][32] (+ LayerNorm wsum [slots][32]); slot = local sub-tile index;
      // GEGLU keeps value and gate columns in two consecutive slots
      constexpr int kSlots = geglu ? 2 * nmine : nmine;
      float* bias_s = reinterpret_cast<float*>(smem + Cfg::kBiasOff) + ew * kWarpFloats;
      float* wsum_s = bias_s + 2 * kSlots * 32;
      const bool has_ln = p.ln_rowstats != nullptr;
      const bool want_stats = !geglu && p.rowstats_out != nullptr;
      // GroupNorm statistics of the output for the GroupNorm that consumes it (north_star: "GroupNorm ... fused into
      // the epilogue"): per sub-tile every thread forms 16 column-pair sums and sums of squares of its row, a
      // transposing butterfly (31 shuffles) reduces them over the warp's 32 rows so that lane i ends up with value i,
      // the values go to shared memory, and once per tile the eight warps combine quadrants and pairs into
      // (row block, micro-group of gn_gran channels) partials in global memory -- fixed order, no atomics.
      const bool want_gn = !geglu && p.gn_part != nullptr;
      float* const stat_s = reinterpret_cast<float*>(smem + Cfg::kStatOff);
      const int et = threadIdx.x - 64;                          // 0..255 within the epilogue group
      const bool per_batch = p.bias_bstride != 0;
      const float* const bias_base0 = p.bias + ((p.bias != nullptr && p.bias_step != nullptr)
                                                    ? static_cast<size_t>(*p.bias_step) * p.bias_step_stride : 0);
      for (int w = cta; w < total_work; w += ncta, ++tl) {
        const WorkItem wi = decode_work(p, w);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        const int h = (hw + static_cast<int>(tl)) & 1;         // my sub-tiles: h, h + 2, ...
        const int grp = (p.dual && wi.mt >= p.mt_single) ? 1 : 0;               // dual launch: second problem's pointers
        const float* const bias_base = grp ? p.bias2 : bias_base0;
        __half* const outp = reinterpret_cast<__half*>(grp ? p.out2 : p.out);
        const __half* const resp = grp ? p.res2 : p.res;
        float* const gnp = grp ? p.gn_part2 : p.gn_part;
        const int m0 = ((wi.mt - grp * p.mt_single) * CG + static_cast<int>(rank)) * kBM + q * 32;   // first row of this warp
        const int m = m0 + lane;
        const bool row_ok = m < p.M;
        int n0 = wi.nt * out_bn;                               // first output column of this tile
        int nb0 = wi.nt * BN;                                  // first bias / weight column of this tile
        size_t orow = static_cast<size_t>(m);                  // output row of this thread
        int gn_blk_mul = 1, gn_blk_add = 0;
        if (!geglu && p.up_shift >= 0) {                       // SEG_UP2x2: scatter to the high-resolution image
          const int par = wi.nt >> p.up_shift;
          n0 -= par * p.up_cout;
          nb0 = n0;
          const int ww = m & ((1 << p.w_shift) - 1), hh = (m >> p.w_shift) & ((1 << p.h_shift) - 1);
          const int bb = m >> (p.w_shift + p.h_shift);
          orow = ((static_cast<size_t>(bb) << (p.h_shift + 1)) + 2 * hh + (par >> 1)) * (2u << p.w_shift) + 2 * ww + (par & 1);
          gn_blk_mul = 4;                                      // statistics: row block = (low-res tile, parity)
          gn_blk_add = par;
        }
        // bias of my sub-tiles -> smem (two batch rows: the 32 rows may straddle a batch boundary); bias and the first
        // residual sub-tile are requested before the accumulator wait (latency overlaps the MMAs)
        int m_last = m0 + 31;
        if (m_last >= p.M) m_last = p.M - 1;
        // a warp whose rows all lie beyond M (M < 128) must not index the per-batch bias table out of range
        const int m_first = m0 < p.M ? m0 : p.M - 1;
        const int b_first = per_batch ? batch_of_row(p, m_first) : 0;
        const int b_last = per_batch ? batch_of_row(p, m_last) : b_first;
        if (has_bias || has_ln) {
          __syncwarp();                                        // previous tile's bias reads are done
#pragma unroll
          for (int s = 0; s < kSlots; ++s) {
            // tile column of slot s, lane: plain -> sub-tile h + 2 s; GEGLU -> value (s even) / gate (s odd) column
            const int jj = geglu ? (s >> 1) : s;
            const int j = h + 2 * jj;
            const int col = (geglu && (s & 1) ? BN / 2 : 0) + j * 32 + lane;
            const int n = nb0 + col;
            const bool ok = j < nsub && wi.nt * BN + col < p.N;
            if (has_bias) {
              bias_s[s * 32 + lane] = ok ? __ldg(bias_base + static_cast<size_t>(b_first) * p.bias_bstride + n) : 0.f;
              bias_s[(kSlots + s) * 32 + lane] = ok ? __ldg(bias_base + static_cast<size_t>(b_last) * p.bias_bstride + n) : 0.f;
            }
            if (has_ln) wsum_s[s * 32 + lane] = ok ? __ldg(p.ln_wsum + n) : 0.f;
          }
          __syncwarp();
        }
        // LayerNorm fold: this row's mean / rstd from the producer's per-N-tile partial sums (fixed order)
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (has_ln && row_ok) {
          float s1 = 0.f, s2 = 0.f;
          const float2* rs = reinterpret_cast<const float2*>(p.ln_rowstats) + static_cast<size_t>(m) * p.ln_parts;
          for (int i = 0; i < p.ln_parts; ++i) {
            const float2 t = __ldg(rs + i);
            s1 += t.x;
            s2 += t.y;
          }
          ln_mean = s1 * p.ln_inv_c;
          ln_rstd = rsqrtf(fmaxf(s2 * p.ln_inv_c - ln_mean * ln_mean, 0.f) + p.ln_eps);
        }
        float st_sum = 0.f, st_sq = 0.f;                       // row statistics of my output columns
        uint32_t rnext[16];
        const __half* const res_row = resp + static_cast<size_t>(m) * p.ldr + n0;
        if (has_res && row_ok && h < nsub) {
          ldg256(res_row + h * 32, rnext);
          ldg256(res_row + h * 32 + 16, rnext + 8);
        }
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        if (tl == 0 && leader) GEMM_TRACE(5);
        if (tl == 1 && leader) GEMM_TRACE(11);
        const int my_b = per_batch ? batch_of_row(p, m) : 0;
        const float* my_bias = bias_s + ((my_b > b_first) ? kSlots * 32 : 0);
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        int jlast = h;                                         // my last sub-tile (releases the accumulator)
        while (jlast + 2 < nsub) jlast += 2;
        if (h >= nsub) {                                       // no work this tile (BN = 32): just release
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
        }
#pragma unroll 1
        for (int j = h, jj = 0; j < nsub; j += 2, ++jj) {
          float v[32];
          uint32_t rcur[16];
          if (tl == 0 && leader && jj == 1) GEMM_TRACE(4);
          if (has_res) {
#pragma unroll
            for (int u = 0; u < 16; ++u) rcur[u] = rnext[u];
            if (j + 2 < nsub && row_ok) {
              ldg256(res_row + (j + 2) * 32, rnext);
              ldg256(res_row + (j + 2) * 32 + 16, rnext + 8);
            }
          }
          if (geglu) {
            float gte[32];
            tmem_ld32(taddr + j * 32, v);
            tmem_ld32(taddr + BN / 2 + j * 32, gte);
            tmem_ld_wait();
            if (tl == 0 && leader && jj == 1) GEMM_TRACE(10);
            if (j == jlast) {                 // my part of the accumulator is read -> release it to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) release_acc(acc);
            }
            const float* wv = wsum_s + (2 * jj) * 32;
            const float* bv = my_bias + (2 * jj) * 32;
            // packed fp32 pairs throughout (64 gate elements per thread and tile).  The bias / LayerNorm options are
            // hoisted OUT of the unrolled loop as compile-time flags of a generic lambda: a warp-uniform branch inside
            // it ends a basic block every four elements, and the scheduler cannot interleave the eight independent
            // rcp -> polynomial -> ex2 chains across those boundaries.
            const f32x2_t nmr = splat_f32x2(-ln_mean * ln_rstd), rs = splat_f32x2(ln_rstd);
            auto geglu_block = [&](auto ln_c, auto bias_c) {
              constexpr bool kLn = decltype(ln_c)::value, kBias = decltype(bias_c)::value;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
                if (kBias) {
                  ba = *reinterpret_cast<const float4*>(bv + i);
                  bg = *reinterpret_cast<const float4*>(bv + 32 + i);
                }
                f32x2_t a0 = pack_f32x2(v[i], v[i + 1]), a1 = pack_f32x2(v[i + 2], v[i + 3]);
                f32x2_t g0 = pack_f32x2(gte[i], gte[i + 1]), g1 = pack_f32x2(gte[i + 2], gte[i + 3]);
                if (kLn) {                                     // rstd * (acc - mean * wsum) + bias
                  const float4 wa = *reinterpret_cast<const float4*>(wv + i);
                  const float4 wg = *reinterpret_cast<const float4*>(wv + 32 + i);
                  a0 = fma_f32x2(a0, rs, fma_f32x2(nmr, pack_f32x2(wa.x, wa.y), pack_f32x2(ba.x, ba.y)));
                  a1 = fma_f32x2(a1, rs, fma_f32x2(nmr, pack_f32x2(wa.z, wa.w), pack_f32x2(ba.z, ba.w)));
                  g0 = fma_f32x2(g0, rs, fma_f32x2(nmr, pack_f32x2(wg.x, wg.y), pack_f32x2(bg.x, bg.y)));
                  g1 = fma_f32x2(g1, rs, fma_f32x2(nmr, pack_f32x2(wg.z, wg.w), pack_f32x2(bg.z, bg.w)));
                } else if (kBias) {
                  a0 = add_f32x2(a0, pack_f32x2(ba.x, ba.y));
                  a1 = add_f32x2(a1, pack_f32x2(ba.z, ba.w));
                  g0 = add_f32x2(g0, pack_f32x2(bg.x, bg.y));
                  g1 = add_f32x2(g1, pack_f32x2(bg.z, bg.w));
                }
                unpack_f32x2(mul_f32x2(a0, gelu_erf_x2(g0)), v[i], v[i + 1]);
                unpack_f32x2(mul_f32x2(a1, gelu_erf_x2(g1)), v[i + 2], v[i + 3]);
              }
            };
            if (has_ln && has_bias) geglu_block(std::true_type{}, std::true_type{});
            else if (has_ln) geglu_block(std::true_type{}, std::false_type{});
            else if (has_bias) geglu_block(std::false_type{}, std::true_type{});
            else geglu_block(std::false_type{}, std::false_type{});
          } else {
            tmem_ld32(taddr + j * 32, v);
            tmem_ld_wait();
            if (tl == 0 && leader && jj == 1) GEMM_TRACE(10);
            if (j == jlast) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) release_acc(acc);
            }
            if (has_ln) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 wq = *reinterpret_cast<const float4*>(wsum_s + jj * 32 + i);
                v[i] = ln_rstd * (v[i] - ln_mean * wq.x);
                v[i + 1] = ln_rstd * (v[i + 1] - ln_mean * wq.y);
                v[i + 2] = ln_rstd * (v[i + 2] - ln_mean * wq.z);
                v[i + 3] = ln_rstd * (v[i + 3] - ln_mean * wq.w);
              }
            }
            if (has_bias) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 bq = *reinterpret_cast<const float4*>(my_bias + jj * 32 + i);
                v[i] += bq.x; v[i + 1] += bq.y; v[i + 2] += bq.z; v[i + 3] += bq.w;
              }
            }
            if (has_res) {
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rcur[u]));
                v[2 * u] += f.x;
                v[2 * u + 1] += f.y;
              }
            }
          }
          if (tl == 0 && leader && jj == 1) GEMM_TRACE(12);
          if (want_stats) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { st_sum += v[i]; st_sq += v[i] * v[i]; }
          }
          if (!geglu && want_gn) {
            float x[32];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const float a = row_ok ? v[2 * u] : 0.f, b = row_ok ? v[2 * u + 1] : 0.f;
              x[u] = a + b;
              x[16 + u] = a * a + b * b;
            }
            // transposing butterfly: after the round with offset `off` a lane keeps the half of its values whose index
            // bit equals its own lane bit, summed over the lane pair -> finally lane i holds sum over rows of x[i]
#pragma unroll
            for (int r = 0; r < 5; ++r) {
              const int off = 16 >> r, n = 32 >> r;
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < n / 2; ++i) {
                const float keep = up ? x[i + n / 2] : x[i];
                const float send = up ? x[i] : x[i + n / 2];
                x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            // lane < 16: sum of pair `lane`; lane >= 16: sum of squares of pair `lane - 16`
            stat_s[((tl & 1) * 4 + q) * BN + (j * 16 + (lane & 15)) * 2 + (lane >> 4)] = x[0];
          }
#if defined(UNIB_WHATIF_EPI) && (UNIB_WHATIF_EPI & 1)
          if (row_ok && p.M < 0) {               // what-if timing aid: the epilogue without its global stores
#else
          if (row_ok) {
#endif
            __half* op = outp + orow * p.ldc + n0 + j * 32;
            uint32_t o[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) o[u] = pack_half2(v[2 * u], v[2 * u + 1]);
            stg256(op, o);
            stg256(op + 16, o + 8);
          }
          if (tl == 0 && leader && jj == 1) GEMM_TRACE(13);
        }
        // two partials per row and N tile (one per warp of the quadrant pair), summed by the consumer in fixed order
        if (want_stats && row_ok)
          *reinterpret_cast<float2*>(p.rowstats_out + (static_cast<size_t>(m) * (2 * p.n_tiles) + 2 * wi.nt + h) * 2) =
              make_float2(st_sum, st_sq);
        if (!geglu && want_gn) {
          named_bar_sync(2, kEpiWarps * 32);                   // the tile's pair partials are complete in stat_s
          const float* sb = stat_s + (tl & 1) * 4 * BN;
          const int ng = BN / p.gn_gran, qpb = p.gn_rows >> 5;  // micro-groups per tile, quadrants per row block
          const int nout = (4 / qpb) * ng * 2;
          const int ngN = p.N / p.gn_gran;
          for (int o = et; o < nout; o += kEpiWarps * 32) {
            const int st = o & 1, mg = (o >> 1) % ng, rb = (o >> 1) / ng;
            float acc = 0.f;
            for (int qq = rb * qpb; qq < (rb + 1) * qpb; ++qq)
              for (int pp = 0; pp < (p.gn_gran >> 1); ++pp) acc += sb[qq * BN + (mg * (p.gn_gran >> 1) + pp) * 2 + st];
            const int row0 = ((wi.mt - grp * p.mt_single) * CG + static_cast<int>(rank)) * kBM + rb * p.gn_rows;
            const int gmg = n0 / p.gn_gran + mg;
            if (row0 < p.M && wi.nt * ng + mg < ngN)
              gnp[(static_cast<size_t>(row0 / p.gn_rows * gn_blk_mul + gn_blk_add) * (gn_blk_mul == 4 ? p.up_cout / p.gn_gran : ngN) + gmg) * 2 + st] = acc;
          }
          // the other parity buffer is written during the next tile; this one is rewritten two tiles from now, after
          // every thread has passed the next tile's barrier
        }
        if (tl == 0 && leader) GEMM_TRACE(9);
      }
      if (leader) GEMM_TRACE(6);
      if (leader) GEMM_TRACE(7);
    } else {
      // ---------------- direct-store epilogue: split-K partials, NCHW outputs, generic shapes ----------------
      for (int w = cta; w < total_work; w += ncta, ++tl) {
        const WorkItem wi = decode_work(p, w);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        const int h = (hw + static_cast<int>(tl)) & 1;
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        if (tl == 0 && leader) GEMM_TRACE(5);
        const int m = (wi.mt * CG + static_cast<int>(rank)) * kBM + q * 32 + lane;
        const bool row_ok = m < p.M;
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        if (p.splits > 1) {
          const int split = w / (p.m_tiles * p.n_tiles);
          float* pp = p.partial + (static_cast<size_t>(split) * p.M + m) * p.N + wi.nt * BN;
#pragma unroll 1
          for (int c = h; c < BN / 32; c += 2) {
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
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) pp[c * 32 + j] = v[j];
              }
            }
          }
        } else {
#pragma unroll 1
          for (int c = h; c < BN / 16; c += 2) {
            float v[16];
            tmem_ld16(taddr + c * 16, v);
            tmem_ld_wait();
            if (row_ok) epilogue_store16(p, v, m, wi.nt * BN + c * 16);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
      }
      if (leader) GEMM_TRACE(7);
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();       // the peer may still be signalling our barriers / the leader's MMAs read our smem
  else __syncthreads();
  // PDL trigger at the END of the CTA's work: a CTA of this kernel owns its SM (shared memory), so an earlier trigger
  // would only park the next kernel's CTAs on SMs the other graph branch could be using
  pdl_launch();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
  if (threadIdx.x == 32) GEMM_TRACE(8);
}

// Split-K finalize: sum fp32 partials over splits in fixed order, then the same epilogue as the fused path.
__global__ void __launch_bounds__(256) gemm_splitk_finalize_kernel(const GemmParams p) {
  pdl_launch();
  pdl_wait();
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
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n + j < p.N) v[j] += pp[j];
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

bool gemm_pair_supported(int bn) { return bn == 128 || bn == 160 || bn == 256; }

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

template <int BN, int CG, int MODE>
static cudaError_t launch_bn(const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, CG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int total_work = p.m_tiles * p.n_tiles * p.splits;     // work items of one CTA (CG = 1) or one CTA pair
  const int slots = num_sms / CG;
  const int grid = CG * (total_work < slots ? total_work : slots);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl_enabled) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, CG, MODE>, maps, p);
  if (e != cudaSuccess) return e;
  static const bool skip_finalize = getenv("UNIB200_SKIP_FINALIZE") != nullptr;   // what-if timing aid (garbage results)
  if (p.splits > 1 && !skip_finalize) {
    const long long total = static_cast<long long>(p.M) * ((p.N + 15) / 16);
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    e = launch_pdl(gemm_splitk_finalize_kernel, dim3(blocks), dim3(256), 0, stream, p);
  }
  return e;
}

template <int BN, int CG>
static cudaError_t launch_mode(const GemmMaps& maps, const GemmParams& p, int num_sms, cudaStream_t stream) {
  switch (p.mode) {
    case MODE_VEC: return launch_bn<BN, CG, MODE_VEC>(maps, p, num_sms, stream);
    case MODE_DIRECT: return launch_bn<BN, CG, MODE_DIRECT>(maps, p, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemm(const GemmMaps& maps, const GemmParams& p, int bn, int num_sms, cudaStream_t stream) {
  if (p.mode == MODE_GEGLU) {                    // GEGLU projections
    if (p.cg == 2) {                             // 256 x 256 pair tiles: the only shape whose operand ingress per FLOP
      if (bn == 256) return launch_bn<256, 2, MODE_GEGLU>(maps, p, num_sms, stream);    // lets the tensor pipe lead
      return cudaErrorInvalidValue;
    }
    switch (bn) {
      case 64: return launch_bn<64, 1, MODE_GEGLU>(maps, p, num_sms, stream);
      case 128: return launch_bn<128, 1, MODE_GEGLU>(maps, p, num_sms, stream);
      case 256: return launch_bn<256, 1, MODE_GEGLU>(maps, p, num_sms, stream);
    }
    return cudaErrorInvalidValue;
  }
  if (p.cg == 2) {
    switch (bn) {
      case 128: return launch_mode<128, 2>(maps, p, num_sms, stream);
      case 160: return launch_mode<160, 2>(maps, p, num_sms, stream);
      case 256: return launch_mode<256, 2>(maps, p, num_sms, stream);
    }
    return cudaErrorInvalidValue;
  }
  switch (bn) {
    case 32: return launch_mode<32, 1>(maps, p, num_sms, stream);
    case 64: return launch_mode<64, 1>(maps, p, num_sms, stream);
    case 128: return launch_mode<128, 1>(maps, p, num_sms, stream);
    case 160: return launch_mode<160, 1>(maps, p, num_sms, stream);
    case 256: return launch_mode<256, 1>(maps, p, num_sms, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace unib
