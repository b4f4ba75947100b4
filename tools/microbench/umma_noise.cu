// Does concurrent softmax-like work slow tcgen05.mma?  One CTA: warp 8 issues bursts (3 x N=128 + 8 x N=48) while
// warps 0-7 run a selectable noise loop.   noise: 0 none, 1 MUFU+FMA, 2 st.shared 16B, 4 tcgen05.ld, combos by OR.
#include "common.cuh"
#include <cstdio>
using namespace unib;

__global__ void __launch_bounds__(320, 1) k(int noise, int bursts, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  const uint32_t base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
  if (warp == 8) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  fence_proxy_async_shared();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, 128);
      const uint32_t idesc_o = make_idesc_f16(128, 48, 0, 1);
      const uint64_t q = make_desc_kmajor_sw128(base), kk = make_desc_kmajor_sw128(base + 32768);
      const uint64_t pd = make_desc_kmajor_sw128(base + 65536), vd = make_desc_mnmajor_sw128(base + 98304, 16384, 1024);
      long long t0 = clock64();
      for (int b = 0; b < bursts; ++b) {
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) umma_f16_ss(tm + (b & 1) * 128, q + 2 * ks, kk + 2 * ks, idesc_s, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_f16_ss(tm + 256 + (b & 1) * 64, pd + ((ks >> 2) * 1024 + (ks & 3) * 2), vd + ks * 128, idesc_o, 1);
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), b & 1);
      }
      long long t1 = clock64();
      out[0] = t1 - t0;
      stop = 1;
    }
  } else if (warp < 8) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
    const uint32_t ta = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    const uint32_t srow = base + 131072 + threadIdx.x * 64;
    float acc = 0.f;
    while (!stop) {
      if (noise & 1) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = fast_exp2(x[i] * 0.5f - 1.0f) + x[(i + 1) & 7];
      }
      if (noise & 2) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(srow + u * 16), "r"(__float_as_uint(x[u])) : "memory");
      }
      if (noise & 4) {
        float v[32];
        tmem_ld32(ta, v);
        tmem_ld_wait();
        acc += v[3];
      }
    }
    sink[threadIdx.x] = x[0] + x[5] + acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; float* sink; long long h[2];
  cudaMalloc(&d, 64); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int noise : {0, 1, 2, 4, 3, 7}) {
    k<<<1, 320, 160 * 1024>>>(noise, 64, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("noise=%d: %.0f cycles per burst (3 x N128 + 8 x N48 + commit + wait)\n", noise, (double)h[0] / 64);
  }
  return 0;
}
