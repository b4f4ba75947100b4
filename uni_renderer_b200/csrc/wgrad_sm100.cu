// Weight-gradient kernel of the implicit-GEMM convolution / linear family for sm_100a (training):
//
//   dW[n, tap, c] = sum over pixels m of  dY[m, n] * X[m (+) tap, c]
//
// i.e. D = dY^T . Xs with the PIXELS as the contraction dimension.  Both operands live in HBM as NHWC matrices
// [pixel][channel], so for the tensor core both are "MN-major" (the contiguous dimension is M / N, not K): the same
// TMA boxes the forward kernel uses for its A operand ({64 channels, 128 pixels}, shifted by the tap for X, zero fill at
// the image border) land in shared memory as [128 pixel rows x 128 B] tiles, and tcgen05.mma reads them through
// MN-major SWIZZLE_128B descriptors with the transpose bits of the instruction descriptor set -- no transpose pass.
// One CTA = one (Cout tile of 128, Cin tile of BN, tap, pixel split): it walks its pixel blocks through a 3-stage ring
// and keeps the 128 x BN fp32 accumulator in TMEM; the epilogue stores it (fp32) either straight into dW or into the
// split's partial slab, summed by wgrad_reduce_kernel in a fixed order (no atomics: reruns are bit-exact).
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
#include "wgrad_sm100.cuh"

#include <cstdlib>

namespace unib {

constexpr int kWgBN = 128;                       // Cin tile
constexpr int kWgStages = 3;
constexpr int kWgABytes = 2 * 128 * 128;         // 128 pixels x 128 Cout channels = two [128 x 64] boxes
constexpr int kWgBBytes = (kWgBN / 64) * 128 * 128;
constexpr int kWgStageBytes = kWgABytes + kWgBBytes;
constexpr int kWgBarOff = kWgStages * kWgStageBytes;
constexpr int kWgSmemBytes = kWgBarOff + 128 + 1024;

__global__ void __launch_bounds__(192, 1)
wgrad_tcgen05_kernel(const __grid_constant__ WgradMaps maps, const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bar_base = base + kWgBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgStages);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kWgBarOff + 8 * (2 * kWgStages + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item: blockIdx.x = ((nt * c_tiles) + ct) * taps + tap ; blockIdx.y = pixel split
  const int tap = blockIdx.x % p.taps;
  const int ct = (blockIdx.x / p.taps) % p.c_tiles;
  const int nt = blockIdx.x / (p.taps * p.c_tiles);
  const int split = blockIdx.y;
  const int kb0 = static_cast<int>((static_cast<long long>(split) * p.m_blocks) / p.splits);
  const int kb1 = static_cast<int>((static_cast<long long>(split + 1) * p.m_blocks) / p.splits);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.dy);
    tma_prefetch_desc(&maps.x);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), kWgBN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer: per pixel block, dY [128 px x 128 Cout] and X shifted by the tap [128 px x BN Cin]
    int dw = 0, dh = 0;
    if (p.taps == 9) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
    uint32_t stage = 0, ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      const int p0 = kb * 128;                                    // first pixel of the block
      const int w0 = p0 & ((1 << p.w_shift) - 1);
      const int h0 = (p0 >> p.w_shift) & ((1 << p.h_shift) - 1);
      const int b0 = p0 >> (p.w_shift + p.h_shift);
      mbar_wait(empty_bar(stage), ph ^ 1);
      if (elect_one()) {
        const uint32_t a_dst = base + stage * kWgStageBytes;
        mbar_arrive_expect_tx(full_bar(stage), kWgStageBytes);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          tma_load_4d(a_dst + c * 16384, &maps.dy, full_bar(stage), nt * 128 + c * 64, w0, h0, b0);
#pragma unroll
        for (int c = 0; c < kWgBN / 64; ++c)
          tma_load_4d(a_dst + kWgABytes + c * 16384, &maps.x, full_bar(stage), ct * kWgBN + c * 64, w0 + dw, h0 + dh, b0);
      }
      if (++stage == kWgStages) { stage = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: D[128 Cout x BN Cin] += dY_blk^T . X_blk, K = 128 pixels per block in steps of 16
    constexpr uint32_t idesc = make_idesc_f16(128, kWgBN, 1, 1);    // both operands MN-major (transposed)
    uint32_t stage = 0, ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(full_bar(stage), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_addr = base + stage * kWgStageBytes;
        // MN-major SWIZZLE_128B: rows = pixels (K), 64-channel blocks 16 KB apart (LBO), 8-pixel atoms 1 KB apart (SBO)
        const uint64_t a_desc = make_desc_mnmajor_sw128(a_addr, 16384, 1024);
        const uint64_t b_desc = make_desc_mnmajor_sw128(a_addr + kWgABytes, 16384, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k)                                // 16 pixels = 2048 B further per step
          umma_f16_ss(tmem_base, a_desc + ((k * 2048) >> 4), b_desc + ((k * 2048) >> 4), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (kb == kb1 - 1) umma_commit(done_bar);
      }
      if (++stage == kWgStages) { stage = 0; ph ^= 1; }
    }
  } else {
    // ---------------- epilogue: TMEM -> fp32 global (thread = one Cout row, 32 Cin columns at a time)
    const int q = warp & 3;
    const int n = nt * 128 + q * 32 + lane;
    float* dst = (p.splits > 1 ? p.partial + static_cast<size_t>(split) * p.N * p.taps * p.C : p.dw) +
                 (static_cast<size_t>(n) * p.taps + tap) * p.C + ct * kWgBN;
    if (kb1 > kb0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
    for (int j = 0; j < kWgBN / 32; ++j) {
      float v[32];
      if (kb1 > kb0) {
        tmem_ld32(taddr + j * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (n < p.N) {
        const int c0 = ct * kWgBN + j * 32;
        if (c0 + 32 <= p.C && (p.C & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(dst + j * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < p.C) dst[j * 32 + i] = v[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgBN);
  }
}

// dW = sum over splits of the partial slabs, fixed order
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* partial, float* dw, long long n, int splits) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += partial[static_cast<size_t>(s) * n + i];
    dw[i] = a;
  }
}

__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t rank) {
  float x;
  const uint32_t remote = mapa_u32(local_addr, rank);
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x) : "r"(remote) : "memory");
  return x;
}

// db[n] = sum over rows of dY[m, n]: one block per 8-channel vector column group, rows strided over threads, fixed-order
// tree in shared memory
__global__ void __launch_bounds__(256) colsum_kernel(const __half* dy, int ld, int M, int N, float* db) {
  __shared__ float red[256][8];
  const int c0 = blockIdx.x * 8;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(dy + static_cast<size_t>(m) * ld + c0);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      s[2 * j] += f.x;
      s[2 * j + 1] += f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = s[j];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[threadIdx.x][j] += red[threadIdx.x + off][j];
    __syncthreads();
  }
  if (threadIdx.x < 8 && c0 + threadIdx.x < N) db[c0 + threadIdx.x] = red[0][threadIdx.x];
}

cudaError_t launch_wgrad(const WgradMaps& maps, const WgradParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(p.n_tiles * p.c_tiles * p.taps, p.splits);
  wgrad_tcgen05_kernel<<<grid, 192, kWgSmemBytes, stream>>>(maps, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (p.splits > 1) {
    const long long n = static_cast<long long>(p.N) * p.taps * p.C;
    int blocks = static_cast<int>((n + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(p.partial, p.dw, n, p.splits);
    e = cudaGetLastError();
  }
  return e;
}

// The same column sums with the rows split over a cluster of 8 CTAs and 32 columns (64 contiguous bytes per row) per
// cluster: four lanes read a row's 64 bytes, 64 row lanes per CTA, four loads in flight per thread; the CTA's partial is
// reduced by a fixed-order tree in shared memory and the cluster's eight partials are summed in rank order by CTA 0
// through distributed shared memory.  (One CTA per 8 columns walked all M rows alone: 40 CTAs and 17 us at N = 320.)
__global__ void __launch_bounds__(256) colsum_cluster_kernel(const __half* dy, int ld, int M, int N, float* db) {
  __shared__ float red[64][32];
  __shared__ float part[32];
  const uint32_t rank = cluster_ctarank(), cs = cluster_nctarank();
  const int c0 = blockIdx.x * 32 + (threadIdx.x & 3) * 8;
  const int rl = threadIdx.x >> 2;                       // 64 row lanes
  const int r0 = static_cast<int>((static_cast<long long>(rank) * M) / cs);
  const int r1 = static_cast<int>((static_cast<long long>(rank + 1) * M) / cs);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  auto acc = [&](const uint4& raw) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      s[2 * j] += f.x;
      s[2 * j + 1] += f.y;
    }
  };
  const __half* src = dy + c0;
  int m = r0 + rl;
  for (; m + 192 < r1; m += 256) {
    const uint4 a0 = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(m) * ld);
    const uint4 a1 = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(m + 64) * ld);
    const uint4 a2 = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(m + 128) * ld);
    const uint4 a3 = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(m + 192) * ld);
    acc(a0); acc(a1); acc(a2); acc(a3);
  }
  for (; m < r1; m += 64) acc(*reinterpret_cast<const uint4*>(src + static_cast<size_t>(m) * ld));
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][(threadIdx.x & 3) * 8 + j] = s[j];
  __syncthreads();
  for (int off = 32; off > 0; off >>= 1) {
    for (int i = threadIdx.x; i < off * 32; i += blockDim.x) red[i >> 5][i & 31] += red[(i >> 5) + off][i & 31];
    __syncthreads();
  }
  if (threadIdx.x < 32) part[threadIdx.x] = red[0][threadIdx.x];
  cluster_sync_all();
  if (rank == 0 && threadIdx.x < 32) {
    const uint32_t la = smem_u32(&part[threadIdx.x]);
    float a = 0.f;
    for (uint32_t r = 0; r < cs; ++r) a += ld_dsmem_f32(la, r);
    db[blockIdx.x * 32 + threadIdx.x] = a;
  }
  cluster_sync_all();                                    // CTA 0 may still be reading the peers' partials
}

// ---------------------------------------------------------------------------------------------------------------
// Training glue between the fp32 master weights (reference layout [O][I][taps], taps = 1 | 9) and the kernels:
// * pack_master_kernel<false>: the forward GEMM's K-major fp16 operand [O][taps * Ipad]  (channels padded to 64 per tap)
// * pack_master_kernel<true>:  the data-gradient GEMM's operand [I][taps * Opad] = the transposed, tap-flipped weights
//   (dX = conv(dY, W'), W'[i][o][ky][kx] = W[o][i][2 - ky][2 - kx])
//   -- one pass over the master each instead of torch's permute / flip / pad / cast chain; the padding columns are
//   zeroed once when the buffer is allocated.  Same round-to-nearest-even cast: bit-identical operands.
// * wgrad_scatter_add_kernel: grad[o][i][t] += dw[o][t][i] -- the weight-gradient kernel's tap-major fp32 output added
//   into the reference-layout flat gradient (was a strided torch permute copy + an add).
// ---------------------------------------------------------------------------------------------------------------
template <bool DGRAD>
__global__ void __launch_bounds__(256) pack_master_kernel(const float* __restrict__ w, __half* __restrict__ out, int O, int I,
                                                          int taps, int pad) {
  const long long total = static_cast<long long>(O) * I;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    int o, i;
    if (DGRAD) { i = static_cast<int>(idx / O); o = static_cast<int>(idx - static_cast<long long>(i) * O); }
    else { o = static_cast<int>(idx / I); i = static_cast<int>(idx - static_cast<long long>(o) * I); }
    const float* src = w + (static_cast<size_t>(o) * I + i) * taps;
    if (DGRAD) {
      __half* dst = out + static_cast<size_t>(i) * taps * pad + o;           // pad = Opad
      for (int t = 0; t < taps; ++t) dst[static_cast<size_t>(taps - 1 - t) * pad] = __float2half_rn(src[t]);
    } else {
      __half* dst = out + static_cast<size_t>(o) * taps * pad + i;           // pad = Ipad
      for (int t = 0; t < taps; ++t) dst[static_cast<size_t>(t) * pad] = __float2half_rn(src[t]);
    }
  }
}
__global__ void __launch_bounds__(256) wgrad_scatter_add_kernel(const float* __restrict__ dw, float* __restrict__ grad, int N,
                                                                int taps, int C) {
  const long long total = static_cast<long long>(N) * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / C), c = static_cast<int>(idx - static_cast<long long>(n) * C);
    float* g = grad + idx * taps;
    const float* d = dw + static_cast<size_t>(n) * taps * C + c;
    for (int t = 0; t < taps; ++t) g[t] += d[static_cast<size_t>(t) * C];
  }
}
// taps == 9: a CTA transposes one (n, 128-channel) slab through shared memory, so both the tap-major reads and the
// reference-layout read-modify-writes are contiguous (the per-thread 9-float walk above runs at a quarter of the
// rate of torch's permute + add)
__global__ void __launch_bounds__(128) wgrad_scatter_add9_kernel(const float* __restrict__ dw, float* __restrict__ grad, int N,
                                                                 int C) {
  __shared__ float tile[9][129];
  const int n = blockIdx.y, c0 = blockIdx.x * 128;
  const int cn = C - c0 < 128 ? C - c0 : 128;
  if (static_cast<int>(threadIdx.x) < cn) {
    const float* d = dw + static_cast<size_t>(n) * 9 * C + c0 + threadIdx.x;
#pragma unroll
    for (int t = 0; t < 9; ++t) tile[t][threadIdx.x] = d[static_cast<size_t>(t) * C];
  }
  __syncthreads();
  float* base = grad + (static_cast<size_t>(n) * C + c0) * 9;
  for (int e = threadIdx.x; e < cn * 9; e += 128) {
    const int c = e / 9, t = e - c * 9;
    base[e] += tile[t][c];
  }
}
// data-gradient operand through a shared-memory transpose: a CTA takes 128 output channels x IT input channels
// (IT * taps <= 72 master floats per output channel, read contiguously) and writes, for each (input channel, tap),
// 128 consecutive fp16 = 256 B of the [I][taps * Opad] operand (the one-thread-per-element kernel above reads its
// 36 bytes at a 46 KB stride: 63 us for a 1280 x 1280 3x3 layer)
__global__ void __launch_bounds__(256) pack_master_dgrad_tile_kernel(const float* __restrict__ w, __half* __restrict__ out,
                                                                     int O, int I, int taps, int opad, int it) {
  __shared__ float tile[128][73];
  const int o0 = blockIdx.x * 128, i0 = blockIdx.y * it;
  const int in = I - i0 < it ? I - i0 : it;              // input channels of this tile
  const int kt = in * taps;                              // contiguous master floats per output channel
  for (int e = threadIdx.x; e < 128 * kt; e += 256) {
    const int ol = e / kt, k = e - ol * kt;
    if (o0 + ol < O) tile[ol][k] = w[(static_cast<size_t>(o0 + ol) * I + i0) * taps + k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < kt * 128; e += 256) {
    const int k = e >> 7, ol = e & 127;
    const int il = k / taps, t = k - il * taps;
    if (o0 + ol < O)
      out[static_cast<size_t>(i0 + il) * taps * opad + static_cast<size_t>(taps - 1 - t) * opad + o0 + ol] =
          __float2half_rn(tile[ol][k]);
  }
}
cudaError_t launch_pack_master(const float* w, __half* out, int O, int I, int taps, int dgrad, cudaStream_t stream) {
  if (dgrad && (taps == 9 || taps == 1) && O >= 128) {
    const int it = taps == 9 ? 8 : 64;
    const dim3 grid((O + 127) / 128, (I + it - 1) / it);
    if (grid.y <= 65535) {
      pack_master_dgrad_tile_kernel<<<grid, 256, 0, stream>>>(w, out, O, I, taps, (O + 63) / 64 * 64, it);
      return cudaGetLastError();
    }
  }
  const long long total = static_cast<long long>(O) * I;
  int blocks = static_cast<int>((total + 255) / 256 > 4736 ? 4736 : (total + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (dgrad) pack_master_kernel<true><<<blocks, 256, 0, stream>>>(w, out, O, I, taps, (O + 63) / 64 * 64);
  else pack_master_kernel<false><<<blocks, 256, 0, stream>>>(w, out, O, I, taps, (I + 63) / 64 * 64);
  return cudaGetLastError();
}
cudaError_t launch_wgrad_scatter_add(const float* dw, float* grad, int N, int taps, int C, cudaStream_t stream) {
  if (taps == 9 && N <= 65535) {
    wgrad_scatter_add9_kernel<<<dim3((C + 127) / 128, N), 128, 0, stream>>>(dw, grad, N, C);
    return cudaGetLastError();
  }
  const long long total = static_cast<long long>(N) * C;
  int blocks = static_cast<int>((total + 255) / 256 > 4736 ? 4736 : (total + 255) / 256);
  if (blocks < 1) blocks = 1;
  wgrad_scatter_add_kernel<<<blocks, 256, 0, stream>>>(dw, grad, N, taps, C);
  return cudaGetLastError();
}

cudaError_t launch_colsum(const __half* dy, int ld, int M, int N, float* db, cudaStream_t stream) {
  if (N % 8) return cudaErrorInvalidValue;
  static const bool old_kernel = getenv("UNIB200_COLSUM_OLD") != nullptr;        // A/B
  if (!old_kernel && N % 32 == 0 && M >= 512 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(N / 32, 8);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 8; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, colsum_cluster_kernel, dy, ld, M, N, db);
  }
  colsum_kernel<<<N / 8, 256, 0, stream>>>(dy, ld, M, N, db);
  return cudaGetLastError();
}

int wgrad_cin_tile() { return kWgBN; }

cudaError_t launch_sum_slabs(const float* partial, float* out, long long n, int slabs, cudaStream_t stream) {
  int blocks = static_cast<int>((n + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(partial, out, n, slabs);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm (+SiLU) backward, one CTA per (group, sample).  Forward: xhat = (x - mu) * rstd, y = gamma * xhat + beta,
// z = silu(y).  Given dz:  dy = dz * silu'(y);  dgamma_c = sum dy * xhat, dbeta_c = sum dy (per-sample partials);
// dxhat = dy * gamma;  dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)) over the group's elements.
// Three sweeps over the group's [HW x cpg] slab (statistics, sums, dx); thread = (channel j, row lane) so the
// per-channel sums need no atomics and every reduction runs in a fixed order.  Correctness-first (the slab is re-read
// from L2); the forward's fused-statistics machinery would serve the first sweep the same way.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) gn_backward_kernel(GnBwdParams p) {
  __shared__ float red[8];
  extern __shared__ float chan[];                 // [2][R][cpg] per-channel partials of the row lanes
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = p.C / p.G;
  const int R = blockDim.x / cpg;                 // row lanes (threads beyond R * cpg idle)
  const int j = threadIdx.x % cpg, rl = threadIdx.x / cpg;
  const bool active = rl < R;
  const int c = g * cpg + j;
  const __half* x = p.x + static_cast<size_t>(b) * p.HW * p.ldx + c;
  const __half* dz = p.dz + static_cast<size_t>(b) * p.HW * p.ldz + c;
  const float n_inv = 1.0f / (static_cast<float>(cpg) * p.HW);
  float s = 0.f, ss = 0.f;
  if (active)
    for (int r = rl; r < p.HW; r += R) {
      const float v = __half2float(x[static_cast<size_t>(r) * p.ldx]);
      s += v;
      ss += v * v;
    }
  const float mu = block_sum_256(s, red) * n_inv;
  const float var = fmaxf(block_sum_256(ss, red) * n_inv - mu * mu, 0.f);
  const float rstd = rsqrtf(var + p.eps);
  const float ga = active ? p.gamma[c] : 0.f, be = active ? p.beta[c] : 0.f;
  float s1 = 0.f, s2 = 0.f, dg = 0.f, db = 0.f;
  if (active)
    for (int r = rl; r < p.HW; r += R) {
      const float xh = (__half2float(x[static_cast<size_t>(r) * p.ldx]) - mu) * rstd;
      float dy = __half2float(dz[static_cast<size_t>(r) * p.ldz]);
      if (p.silu) {
        const float y = ga * xh + be;
        const float sg = 1.0f / (1.0f + __expf(-y));
        dy *= sg * (1.0f + y * (1.0f - sg));
      }
      dg += dy * xh;
      db += dy;
      const float dxh = dy * ga;
      s1 += dxh;
      s2 += dxh * xh;
    }
  const float m1 = block_sum_256(s1, red) * n_inv;
  const float m2 = block_sum_256(s2, red) * n_inv;
  if (active) {
    chan[rl * cpg + j] = dg;
    chan[(R + rl) * cpg + j] = db;
  }
  __syncthreads();
  if (threadIdx.x < cpg) {
    float a = 0.f, bb = 0.f;
    for (int r = 0; r < R; ++r) { a += chan[r * cpg + threadIdx.x]; bb += chan[(R + r) * cpg + threadIdx.x]; }
    p.dgamma_part[static_cast<size_t>(b) * p.C + g * cpg + threadIdx.x] = a;
    p.dbeta_part[static_cast<size_t>(b) * p.C + g * cpg + threadIdx.x] = bb;
  }
  if (active) {
    __half* dx = p.dx + static_cast<size_t>(b) * p.HW * p.lddx + c;
    for (int r = rl; r < p.HW; r += R) {
      const float xh = (__half2float(x[static_cast<size_t>(r) * p.ldx]) - mu) * rstd;
      float dy = __half2float(dz[static_cast<size_t>(r) * p.ldz]);
      if (p.silu) {
        const float y = ga * xh + be;
        const float sg = 1.0f / (1.0f + __expf(-y));
        dy *= sg * (1.0f + y * (1.0f - sg));
      }
      dx[static_cast<size_t>(r) * p.lddx] = __float2half_rn(rstd * (dy * ga - m1 - xh * m2));
    }
  }
}

// The same backward as a thread-block CLUSTER per sample (up to 16 CTAs, like the forward's gn_cluster_kernel): a thread
// owns one 8-channel vector column (128-bit loads, rows in parallel across the CTA) instead of one channel of one group
// (2-byte loads, 20 contiguous bytes per row at C / G = 10), per-channel sums are reduced over the CTA in shared memory
// and over the cluster through distributed shared memory in a fixed order.  Sweep 1: per-group sum / sum of squares ->
// mu, rstd.  Sweep 2: per-channel dgamma = sum dy xhat, dbeta = sum dy; the group means of dxhat and dxhat * xhat follow
// from them (sum_c gamma_c dbeta_c, sum_c gamma_c dgamma_c).  Sweep 3: dx.  85 -> ~20 us on a 64x64 x 320 sample batch.

__global__ void __launch_bounds__(512) gn_backward_cluster_kernel(GnBwdParams p) {
  extern __shared__ float sm[];                    // red[2][rpb][C] | chs[2][C] (this CTA's per-channel sums)
  __shared__ float part[2 * 64];                   // this CTA's per-group partials -- read by the peers
  __shared__ float gstat[4 * 64];                  // (mu, rstd, m1, m2) per group
  const int C = p.C, CV = C >> 3, cpg = C / p.G;
  const int rpb = blockDim.x / CV;
  const int b = blockIdx.y;
  const uint32_t rank = cluster_ctarank(), cs = cluster_nctarank();
  const int r0 = static_cast<int>((static_cast<long long>(rank) * p.HW) / cs);
  const int r1 = static_cast<int>((static_cast<long long>(rank + 1) * p.HW) / cs);
  float* red0 = sm;
  float* red1 = sm + rpb * C;
  float* chs0 = sm + 2 * rpb * C;
  float* chs1 = chs0 + C;
  const int v = threadIdx.x % CV, rsub = threadIdx.x / CV;
  const bool active = rsub < rpb;
  const int c0 = v * 8;
  const __half* x = p.x + static_cast<size_t>(b) * p.HW * p.ldx + c0;
  const __half* dz = p.dz + static_cast<size_t>(b) * p.HW * p.ldz + c0;
  const float n_inv = 1.0f / (static_cast<float>(cpg) * p.HW);
  auto unpack8 = [](const uint4& raw, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
  };
  // CTA-wide column sums of two per-thread accumulators -> chs0 / chs1; group sums (optionally gamma-weighted) -> part
  auto reduce_cta = [&](const float* a, const float* q, bool weighted) {
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red0[rsub * C + c0 + j] = a[j];
        red1[rsub * C + c0 + j] = q[j];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < rpb; ++r) { s0 += red0[r * C + c]; s1 += red1[r * C + c]; }
      chs0[c] = s0;
      chs1[c] = s1;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * p.G; t += blockDim.x) {
      const int st = t & 1, g = t >> 1;
      const float* src = st ? chs1 : chs0;
      float acc = 0.f;
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) acc += weighted ? src[c] * p.gamma[c] : src[c];
      part[2 * g + st] = acc;
    }
  };
  // sum of every CTA's group partial, fixed order
  auto cluster_sum = [&](int idx) {
    const uint32_t la = smem_u32(&part[idx]);
    float acc = 0.f;
    for (uint32_t r = 0; r < cs; ++r) acc += ld_dsmem_f32(la, r);
    return acc;
  };

  // ---- sweep 1: statistics
  float a[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = 0.f; q[j] = 0.f; }
  if (active) {
    int r = r0 + rsub;
    for (; r + rpb < r1; r += 2 * rpb) {
      const uint4 u0 = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * p.ldx);
      const uint4 u1 = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(r + rpb) * p.ldx);
      float f0[8], f1[8];
      unpack8(u0, f0);
      unpack8(u1, f1);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] += f0[j] + f1[j]; q[j] += f0[j] * f0[j] + f1[j] * f1[j]; }
    }
    if (r < r1) {
      float f0[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * p.ldx), f0);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] += f0[j]; q[j] += f0[j] * f0[j]; }
    }
  }
  reduce_cta(a, q, false);
  cluster_sync_all();
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    const float mu = cluster_sum(2 * g) * n_inv;
    const float var = fmaxf(cluster_sum(2 * g + 1) * n_inv - mu * mu, 0.f);
    gstat[4 * g] = mu;
    gstat[4 * g + 1] = rsqrtf(var + p.eps);
  }
  cluster_sync_all();                              // every peer has read `part`; gstat is visible CTA-wide
  // ---- sweep 2: per-channel dgamma / dbeta
  float mu[8], rstd[8], ga[8], be[8];
  int grp[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    grp[j] = (c0 + j) / cpg;
    mu[j] = active ? gstat[4 * grp[j]] : 0.f;
    rstd[j] = active ? gstat[4 * grp[j] + 1] : 0.f;
    ga[j] = active ? p.gamma[c0 + j] : 0.f;
    be[j] = active ? p.beta[c0 + j] : 0.f;
    a[j] = 0.f;                                    // dgamma
    q[j] = 0.f;                                    // dbeta
  }
  auto dy_of = [&](float xh, float d, int j) {
    if (p.silu) {
      const float y = ga[j] * xh + be[j];
      const float sg = 1.0f / (1.0f + __expf(-y));
      d *= sg * (1.0f + y * (1.0f - sg));
    }
    return d;
  };
  if (active) {
    for (int r = r0 + rsub; r < r1; r += rpb) {
      float fx[8], fd[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * p.ldx), fx);
      unpack8(*reinterpret_cast<const uint4*>(dz + static_cast<size_t>(r) * p.ldz), fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (fx[j] - mu[j]) * rstd[j];
        const float d = dy_of(xh, fd[j], j);
        a[j] += d * xh;
        q[j] += d;
      }
    }
  }
  reduce_cta(a, q, true);                          // part[2g] = sum_c gamma dgamma, part[2g + 1] = sum_c gamma dbeta
  cluster_sync_all();
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    gstat[4 * g + 3] = cluster_sum(2 * g) * n_inv;        // m2 = mean(dxhat * xhat)
    gstat[4 * g + 2] = cluster_sum(2 * g + 1) * n_inv;    // m1 = mean(dxhat)
  }
  // per-sample dgamma / dbeta: CTA `rank` sums the channels [rank * C / cs, (rank + 1) * C / cs) over the cluster
  {
    const int cb = static_cast<int>((static_cast<long long>(rank) * C) / cs);
    const int ce = static_cast<int>((static_cast<long long>(rank + 1) * C) / cs);
    for (int c = cb + threadIdx.x; c < ce; c += blockDim.x) {
      float dg = 0.f, db = 0.f;
      const uint32_t l0 = smem_u32(&chs0[c]), l1 = smem_u32(&chs1[c]);
      for (uint32_t r = 0; r < cs; ++r) { dg += ld_dsmem_f32(l0, r); db += ld_dsmem_f32(l1, r); }
      p.dgamma_part[static_cast<size_t>(b) * C + c] = dg;
      p.dbeta_part[static_cast<size_t>(b) * C + c] = db;
    }
  }
  __syncthreads();
  // ---- sweep 3: dx = rstd * (dy gamma - m1 - xhat m2)
  if (active) {
    float m1[8], m2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m1[j] = gstat[4 * grp[j] + 2]; m2[j] = gstat[4 * grp[j] + 3]; }
    __half* dx = p.dx + static_cast<size_t>(b) * p.HW * p.lddx + c0;
    for (int r = r0 + rsub; r < r1; r += rpb) {
      float fx[8], fd[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * p.ldx), fx);
      unpack8(*reinterpret_cast<const uint4*>(dz + static_cast<size_t>(r) * p.ldz), fd);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float out2[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int jj = 2 * j + k;
          const float xh = (fx[jj] - mu[jj]) * rstd[jj];
          const float d = dy_of(xh, fd[jj], jj);
          out2[k] = rstd[jj] * (d * ga[jj] - m1[jj] - xh * m2[jj]);
        }
        o[j] = pack_half2(out2[0], out2[1]);
      }
      *reinterpret_cast<uint4*>(dx + static_cast<size_t>(r) * p.lddx) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  cluster_sync_all();                              // peers may still be reading this CTA's shared memory
}

static int gn_bwd_cluster_size(int HW, int max_cs) {
  int cs = 1;
  while (cs * 2 <= max_cs && HW / (cs * 2) >= 8) cs *= 2;
  return cs;
}

cudaError_t launch_gn_backward(const GnBwdParams& p, int B, cudaStream_t stream) {
  if (p.C % p.G || p.C / p.G > 256) return cudaErrorInvalidValue;
  const int CV = p.C >> 3;
  auto al16 = [](const void* q, int ld) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && ld % 8 == 0; };
  const bool vec_ok = p.C % 8 == 0 && CV <= 512 && p.G <= 64 && al16(p.x, p.ldx) && al16(p.dz, p.ldz) && al16(p.dx, p.lddx);
  static const bool old_kernel = getenv("UNIB200_GN_BWD_OLD") != nullptr;       // A/B: one CTA per (group, sample)
  if (vec_ok && !old_kernel) {
    const int rpb = 512 / CV;
    const size_t smem = (static_cast<size_t>(2) * rpb * p.C + 2 * p.C) * sizeof(float);
    static int max_cs = 0;
    if (max_cs == 0) {
      max_cs = 8;
      if (cudaFuncSetAttribute(gn_backward_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
          cudaFuncSetAttribute(gn_backward_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) == cudaSuccess) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3(16, 1);
        q.blockDim = dim3(512);
        q.dynamicSmemBytes = 96 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        q.attrs = at;
        q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, gn_backward_cluster_kernel, &q) == cudaSuccess && n > 0) max_cs = 16;
      }
      cudaGetLastError();
    }
    if (smem <= 96 * 1024) {
      const int cs = gn_bwd_cluster_size(p.HW, max_cs);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs, B);
      cfg.blockDim = dim3(512);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, gn_backward_cluster_kernel, p);
    }
  }
  const int cpg = p.C / p.G, R = 256 / cpg;
  gn_backward_kernel<<<dim3(p.G, B), 256, 2 * R * cpg * sizeof(float), stream>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Transformer-block backward helpers (row-wise, one warp per row; correctness-first)
// ---------------------------------------------------------------------------------------------------------------
// LayerNorm backward: xhat = (x - mu) rstd, y = gamma xhat + beta.  dx = rstd (dxh - mean(dxh) - xhat mean(dxh xhat)),
// dxh = dy gamma.  dgamma / dbeta: every warp adds its rows into a per-CTA slab [gridDim.x][2][C] (fixed order inside the
// CTA: warps take rows round-robin and the 8 warp partials are summed in order), summed over CTAs by sum_slabs.
__global__ void __launch_bounds__(256) layernorm_backward_kernel(const __half* x, const __half* dy, __half* dx,
                                                                 const float* gamma, float* slabs, int rows, int C,
                                                                 float eps) {
  extern __shared__ float wsm[];                   // [8 warps][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* mine = wsm + warp * 2 * C;
  for (int c = lane; c < 2 * C; c += 32) mine[c] = 0.f;
  __syncwarp();
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const __half* xr = x + static_cast<size_t>(r) * C;
    const __half* dr = dy + static_cast<size_t>(r) * C;
    float s = 0.f, ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = __half2float(xr[c]);
      s += v;
      ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const float mu = s / C;
    const float rstd = rsqrtf(fmaxf(ss / C - mu * mu, 0.f) + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xh = (__half2float(xr[c]) - mu) * rstd;
      const float d = __half2float(dr[c]);
      mine[c] += d * xh;
      mine[C + c] += d;
      const float dxh = d * gamma[c];
      s1 += dxh;
      s2 += dxh * xh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / C, m2 = s2 / C;
    for (int c = lane; c < C; c += 32) {
      const float xh = (__half2float(xr[c]) - mu) * rstd;
      dx[static_cast<size_t>(r) * C + c] = __float2half_rn(rstd * (__half2float(dr[c]) * gamma[c] - m1 - xh * m2));
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += wsm[w * 2 * C + c];
    slabs[static_cast<size_t>(blockIdx.x) * 2 * C + c] = a;
  }
}

cudaError_t launch_layernorm_backward(const __half* x, const __half* dy, __half* dx, const float* gamma, float* dgamma,
                                      float* dbeta, float* slabs, int max_slabs, int rows, int C, float eps,
                                      cudaStream_t stream) {
  int blocks = (rows + 7) / 8;
  if (blocks > max_slabs) blocks = max_slabs;
  if (blocks > 592) blocks = 592;
  if (blocks < 1 || static_cast<size_t>(8) * 2 * C * 4 > 96 * 1024) return cudaErrorInvalidValue;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(layernorm_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  layernorm_backward_kernel<<<blocks, 256, 8 * 2 * C * sizeof(float), stream>>>(x, dy, dx, gamma, slabs, rows, C, eps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // slabs are [blocks][2][C]: dgamma = sum over blocks of row 0, dbeta of row 1 -> sum the [2C] vectors, then split
  e = launch_sum_slabs(slabs, dgamma, 2 * C, blocks, stream);      // dgamma buffer must hold 2*C floats: [dgamma | dbeta]
  (void)dbeta;
  return e;
}

// GEGLU: h = a * gelu(g) with [a | g] = proj[:, :inner | inner:]  (diffusers GEGLU, exact erf GELU)
__global__ void __launch_bounds__(256) geglu_forward_kernel(const __half* proj, __half* out, long long rows, int inner) {
  const long long n = rows * inner;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / inner;
    const int c = static_cast<int>(i - r * inner);
    const float a = __half2float(proj[r * 2 * inner + c]), g = __half2float(proj[r * 2 * inner + inner + c]);
    out[i] = __float2half_rn(a * 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)));
  }
}
__global__ void __launch_bounds__(256) geglu_backward_kernel(const __half* proj, const __half* dout, __half* dproj,
                                                             long long rows, int inner) {
  const long long n = rows * inner;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / inner;
    const int c = static_cast<int>(i - r * inner);
    const float a = __half2float(proj[r * 2 * inner + c]), g = __half2float(proj[r * 2 * inner + inner + c]);
    const float d = __half2float(dout[i]);
    const float cdf = 0.5f * (1.0f + erff(g * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * g * g);
    dproj[r * 2 * inner + c] = __float2half_rn(d * g * cdf);                       // d/da = gelu(g)
    dproj[r * 2 * inner + inner + c] = __float2half_rn(d * a * (cdf + g * pdf));    // d/dg = a * gelu'(g)
  }
}
cudaError_t launch_geglu(const __half* proj, const __half* dout, __half* out, long long rows, int inner, int backward,
                         cudaStream_t stream) {
  long long n = rows * inner;
  int blocks = static_cast<int>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
  if (backward) geglu_backward_kernel<<<blocks, 256, 0, stream>>>(proj, dout, out, rows, inner);
  else geglu_forward_kernel<<<blocks, 256, 0, stream>>>(proj, out, rows, inner);
  return cudaGetLastError();
}

// softmax backward, one warp per row: dS = scale * P o (dP - sum_j dP_j P_j), written over dP
__global__ void __launch_bounds__(256) softmax_backward_kernel(const __half* P, __half* dP, int rows, int n, int ld,
                                                               float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const __half* pr = P + static_cast<size_t>(r) * ld;
    __half* dr = dP + static_cast<size_t>(r) * ld;
    float s = 0.f;
    for (int c = lane; c < n; c += 32) s += __half2float(pr[c]) * __half2float(dr[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int c = lane; c < n; c += 32)
      dr[c] = __float2half_rn(scale * __half2float(pr[c]) * (__half2float(dr[c]) - s));
  }
}
cudaError_t launch_softmax_backward(const __half* P, __half* dP, int rows, int n, int ld, float scale, cudaStream_t stream) {
  int blocks = (rows + 7) / 8;
  if (blocks > 1184) blocks = 1184;
  softmax_backward_kernel<<<blocks, 256, 0, stream>>>(P, dP, rows, n, ld, scale);
  return cudaGetLastError();
}

// fp32 [rows, cols] (contiguous) -> fp16 matrix with leading dimension ld (a column slice of a wider gradient matrix)
__global__ void __launch_bounds__(256) cvt_f32_f16_kernel(const float* src, __half* dst, long long rows, int cols, int ld) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    dst[r * ld + (i - r * cols)] = __float2half_rn(src[i]);
  }
}
cudaError_t launch_cvt_f32_f16(const float* src, __half* dst, long long rows, int cols, int ld, cudaStream_t stream) {
  long long n = rows * cols;
  int blocks = static_cast<int>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
  cvt_f32_f16_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(src, dst, rows, cols, ld);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// network-level training glue (uni_renderer_b200/trainer.py): SiLU forward / backward on fp16 vectors (time-embedding
// path), the two resampling adjoints and the optimizer.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) silu_f16_kernel(const __half* x, const __half* dy, __half* out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = __half2float(x[i]);
    const float sg = 1.0f / (1.0f + __expf(-v));
    out[i] = __float2half_rn(dy ? __half2float(dy[i]) * sg * (1.0f + v * (1.0f - sg)) : v * sg);
  }
}
cudaError_t launch_silu_f16(const __half* x, const __half* dy, __half* out, long long n, cudaStream_t stream) {
  int blocks = static_cast<int>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
  silu_f16_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(x, dy, out, n);
  return cudaGetLastError();
}

// adjoint of nearest-2x upsampling: dst[b, h, w, :] = sum of the 2x2 block of src [B, 2H, 2W, C] (fp32 sum, 8 channels
// per thread)
__global__ void __launch_bounds__(256) pool2x2_sum_kernel(const __half* src, __half* dst, int B, int H, int W, int C) {
  const int CV = C >> 3;
  const long long n = static_cast<long long>(B) * H * W * CV;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    const long long pix = i / CV;
    const int w = static_cast<int>(pix % W), h = static_cast<int>((pix / W) % H), b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const size_t sp = (static_cast<size_t>(b) * 2 * H + 2 * h + (q >> 1)) * (2 * W) + 2 * w + (q & 1);
      const uint4 raw = *reinterpret_cast<const uint4*>(src + sp * C + cv * 8);
      const __half2* hp = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hp[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack_half2(acc[2 * j], acc[2 * j + 1]);
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(pix) * C + cv * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
cudaError_t launch_pool2x2_sum(const __half* src, __half* dst, int B, int H, int W, int C, cudaStream_t stream) {
  const long long n = static_cast<long long>(B) * H * W * (C >> 3);
  int blocks = static_cast<int>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
  pool2x2_sum_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(src, dst, B, H, W, C);
  return cudaGetLastError();
}

// zero insertion: dst [B, 2H, 2W, C] = src [B, H, W, C] at the even pixels, 0 elsewhere -- turns the data / weight
// gradient of a stride-2 3x3 convolution into the stride-1 kernels' problem (trainer.py)
__global__ void __launch_bounds__(256) scatter2x_kernel(const __half* src, __half* dst, int B, int H, int W, int C) {
  const int CV = C >> 3;
  const long long n = static_cast<long long>(B) * 4 * H * W * CV;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % CV);
    const long long pix = i / CV;
    const int x = static_cast<int>(pix % (2 * W)), y = static_cast<int>((pix / (2 * W)) % (2 * H));
    const int b = static_cast<int>(pix / (4LL * W * H));
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (!(x & 1) && !(y & 1))
      v = *reinterpret_cast<const uint4*>(src + ((static_cast<size_t>(b) * H + (y >> 1)) * W + (x >> 1)) * C + cv * 8);
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(pix) * C + cv * 8) = v;
  }
}
cudaError_t launch_scatter2x(const __half* src, __half* dst, int B, int H, int W, int C, cudaStream_t stream) {
  const long long n = static_cast<long long>(B) * 4 * H * W * (C >> 3);
  int blocks = static_cast<int>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
  scatter2x_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(src, dst, B, H, W, C);
  return cudaGetLastError();
}

// AdamW on flat fp32 buffers, torch.optim.AdamW's arithmetic: p *= 1 - lr wd; m, v moments of g * grad_scale (the
// inverse loss scale and the clip coefficient); p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                          float wd, float bc1, float bc2_sqrt, float grad_scale) {
  const float gi = g * grad_scale;
  const float mi = b1 * m + (1.0f - b1) * gi;
  const float vi = b2 * v + (1.0f - b2) * gi * gi;
  m = mi;
  v = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p = p * (1.0f - lr * wd) - (lr / bc1) * (mi / denom);
}
// One pass over the flat fp32 buffers (p, g, m, v read; p, m, v written: 28 bytes per parameter, 49 GB at SD-1.5 size):
// pure HBM streaming, so every thread keeps two 128-bit loads of each array in flight (the scalar grid-stride version
// ran at 4.4 TB/s).  Same arithmetic per element as before: results are bit-identical.
__global__ void __launch_bounds__(256) adamw_kernel(float* p, const float* g, float* m, float* v, long long n, float lr,
                                                    float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float grad_scale) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  auto upd = [&](float4& pp, const float4& gg, float4& mm, float4& vv) {
    adamw_one(pp.x, gg.x, mm.x, vv.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
    adamw_one(pp.y, gg.y, mm.y, vv.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
    adamw_one(pp.z, gg.z, mm.z, vv.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
    adamw_one(pp.w, gg.w, mm.w, vv.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
  };
  for (; i + stride < n4; i += 2 * stride) {
    float4 pa = p4[i], pb = p4[i + stride];
    const float4 ga = __ldg(g4 + i), gb = __ldg(g4 + i + stride);
    float4 ma = m4[i], mb = m4[i + stride];
    float4 va = v4[i], vb = v4[i + stride];
    upd(pa, ga, ma, va);
    upd(pb, gb, mb, vb);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    p4[i + stride] = pb; m4[i + stride] = mb; v4[i + stride] = vb;
  }
  if (i < n4) {
    float4 pa = p4[i], ma = m4[i], va = v4[i];
    const float4 ga = __ldg(g4 + i);
    upd(pa, ga, ma, va);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
  }
  // tail (n not a multiple of 4)
  for (long long t = (n4 << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < n; t += stride)
    adamw_one(p[t], g[t], m[t], v[t], lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
}
__global__ void __launch_bounds__(256) adamw_scalar_kernel(float* p, const float* g, float* m, float* v, long long n,
                                                           float lr, float b1, float b2, float eps, float wd, float bc1,
                                                           float bc2_sqrt, float grad_scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    adamw_one(p[i], g[i], m[i], v[i], lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
}
cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                         float wd, int step, float grad_scale, cudaStream_t stream) {
  const float bc1 = 1.0f - powf(b1, static_cast<float>(step));
  const float bc2_sqrt = sqrtf(1.0f - powf(b2, static_cast<float>(step)));
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (aligned) {
    const long long work = (n / 4 + 1) / 2 + 1;                // float4 pairs per thread-iteration
    int blocks = static_cast<int>((work + 255) / 256 > 2368 ? 2368 : (work + 255) / 256);
    adamw_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
  } else {
    int blocks = static_cast<int>((n + 255) / 256 > 2368 ? 2368 : (n + 255) / 256);
    adamw_scalar_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt,
                                                                     grad_scale);
  }
  return cudaGetLastError();
}

}  // namespace unib
