// Micro-benchmarks on one SM: tcgen05.mma issue/execution rate vs N, TMEM load rate, MUFU ex2 rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../uni_renderer_b200/csrc umma_rate.cu -o umma_rate
#include "common.cuh"
#include <cstdio>
using namespace unib;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(160, 1) mma_kernel(int N, int nmma, int b_mn_major, long long* out, int ts = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 4) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_shared();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, 0, b_mn_major);
    const uint64_t a_desc = make_desc_kmajor_sw128(base);
    const uint64_t b_desc = b_mn_major ? make_desc_mnmajor_sw128(base + 32768, 16384, 1024) : make_desc_kmajor_sw128(base + 32768);
    long long t0 = clock64();
    if (ts) {
      for (int i = 0; i < nmma; ++i) umma_f16_ts(tm, tm + 256 + 8 * (i & 7), b_desc + (b_mn_major ? 128 * (i & 3) : 2 * (i & 3)), idesc, i > 0);
    } else {
      for (int i = 0; i < nmma; ++i) umma_f16_ss(tm, a_desc + 2 * (i & 3), b_desc + (b_mn_major ? 128 * (i & 3) : 2 * (i & 3)), idesc, i > 0);
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// 4 warps (or 8) each tcgen05.ld 128 columns `iters` times
__global__ void __launch_bounds__(288, 1) tmem_ld_kernel(int nwarps, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 8) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp < nwarps) {
    const uint32_t ta = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float acc = 0.f;
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      float v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(ta + c * 32, v + c * 32);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc += v[i];
    }
    long long t1 = clock64();
    if (lane == 0) out[warp] = t1 - t0;
    sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

__global__ void mufu_kernel(int iters, long long* out, float* sink) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fast_exp2(x[i]) - 1.0f;
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 64 * 8); cudaMalloc(&sink, 4096);
  long long h[16];
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int Ns[] = {16, 32, 48, 64, 96, 128, 160, 256};
  for (int mn = 0; mn < 2; ++mn)
    for (int N : Ns) {
      for (int nmma : {8, 64}) {
        mma_kernel<<<1, 160, 96 * 1024>>>(N, nmma, mn, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma N=%d failed: %s\n", N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("UMMA M=128 N=%3d K=16 b_mn=%d x%2d: issue %5lld cyc (%.1f/mma)  complete %5lld cyc (%.1f/mma)\n", N, mn, nmma, h[0],
               (double)h[0] / nmma, h[1], (double)h[1] / nmma);
      }
    }
  for (int N : {16, 48, 64, 80, 128, 160, 256})
    for (int nmma : {8, 64}) {
      mma_kernel<<<1, 160, 96 * 1024>>>(N, nmma, 1, d, 1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("TS mma N=%d failed: %s\n", N, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("UMMA-TS (A in TMEM) M=128 N=%3d K=16 x%2d: issue %5lld cyc (%.1f/mma)  complete %5lld cyc (%.1f/mma)\n", N, nmma, h[0],
             (double)h[0] / nmma, h[1], (double)h[1] / nmma);
    }
  for (int nw : {1, 4, 8}) {
    tmem_ld_kernel<<<1, 288>>>(nw, 64, d, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("TMEM ld 128 cols x 64 iters, %d warps: %lld cyc/warp -> %.1f cyc per 128-col row-block load, %.1f B/clk/SM\n", nw, h[0],
           (double)h[0] / 64, nw * 32.0 * 128 * 4 * 64 / h[0]);
  }
  for (int nt : {32, 128, 256, 512}) {
    mufu_kernel<<<1, nt>>>(256, d, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("MUFU ex2: %d threads, %lld cyc for %d ex2 -> %.2f ex2/clk/SM\n", nt, h[0], nt * 256 * 8, nt * 256.0 * 8 / h[0]);
  }
  return 0;
}
