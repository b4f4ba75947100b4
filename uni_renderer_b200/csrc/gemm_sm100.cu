// Implicit-GEMM convolution / linear kernel for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma (cta_group::1,
// UMMA 128 x BN x 16, fp16 in / fp32 accumulate in TMEM, double-buffered accumulators) -> tcgen05.ld epilogue.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue.
// Persistent: each CTA walks work items (split, m_tile, n_tile) with stride gridDim.x.
#include "gemm_sm100.cuh"

namespace unib {

template <int BN>
struct GemmCfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment for SWIZZLE_128B");
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

// 16 accumulator columns [n, n+16) of row m -> final output.  `v` already holds fp32 sums.
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

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;          // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + Cfg::kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes + (2 * Cfg::kStages + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kMaxAMaps; ++i) tma_prefetch_desc(&maps.a[i]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
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
    uint32_t tl = 0;
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
        for (int c = 0; c < BN / 16; ++c) {
          float v[16];
          tmem_ld16(taddr + c * 16, v);
          tmem_ld_wait();
          const int n = wi.nt * BN + c * 16;
          if (row_ok) {
            if (n + 16 <= p.N) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(pp + c * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int j = 0; j < 16 && n + j < p.N; ++j) pp[c * 16 + j] = v[j];
            }
          }
        }
      } else if (p.flags & EPI_GEGLU) {
        constexpr int HALF = BN / 2;
        const int nout = p.N / 2;
#pragma unroll 1
        for (int c = 0; c < HALF / 16; ++c) {
          float a[16], g[16];
          tmem_ld16(taddr + c * 16, a);
          tmem_ld16(taddr + HALF + c * 16, g);
          tmem_ld_wait();
          if (row_ok) {
            const float* ba = p.bias + wi.nt * BN + c * 16;
            const float* bg = ba + HALF;
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = (a[j] + ba[j]) * gelu_erf_f(g[j] + bg[j]);
            const int n = wi.nt * HALF + c * 16;
            if (n + 16 <= nout) {
              __half* op = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(m) * p.ldc + n;
              uint4 o0, o1;
              o0.x = pack_half2(o[0], o[1]);   o0.y = pack_half2(o[2], o[3]);
              o0.z = pack_half2(o[4], o[5]);   o0.w = pack_half2(o[6], o[7]);
              o1.x = pack_half2(o[8], o[9]);   o1.y = pack_half2(o[10], o[11]);
              o1.z = pack_half2(o[12], o[13]); o1.w = pack_half2(o[14], o[15]);
              reinterpret_cast<uint4*>(op)[0] = o0;
              reinterpret_cast<uint4*>(op)[1] = o1;
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
  }
  return 0;
}

int gemm_pick_bn(int N, int flags) {
  const int cands[4] = {160, 128, 64, 32};
  for (int i = 0; i < 4; ++i)
    if (N % cands[i] == 0) return cands[i];
  (void)flags;
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
  }
  return cudaErrorInvalidValue;
}

}  // namespace unib
