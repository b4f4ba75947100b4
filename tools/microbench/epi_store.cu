// How fast can one SM's epilogue warps write / read a 128 x 160 fp16 tile (row pitch 640 B) with different access
// shapes?  148 CTAs x 256 threads (8 warps, like the GEMM epilogue); each CTA walks `tiles` tiles.
//   mode 0: thread = row, 2 x 256-bit stores of 64 contiguous bytes per 32-column sub-tile (the current epilogue)
//   mode 1: same bytes, but 4 lanes cover one row's 64 B with 128-bit stores (8 rows per instruction)
//   mode 2: 8 lanes cover one row's 128 B (64 columns) with 128-bit stores: full 128 B lines (4 rows per instruction)
//   mode 3: mode 0 + residual read (2 x 256-bit loads per sub-tile) ; mode 4: mode 2 + coalesced residual read
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 epi_store.cu -o epi_store
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
constexpr int N = 320, BN = 160;
__global__ void __launch_bounds__(256) k(__half* out, const __half* res, int tiles_per_cta, int mode, long long* cyc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, h = warp >> 2;
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int tile = blockIdx.x * tiles_per_cta + t;
    const int m0 = (tile >> 1) * 128 + q * 32, n0 = (tile & 1) * BN;
    if (mode == 0 || mode == 3) {
      const int m = m0 + lane;
      for (int j = h; j < BN / 32; j += 2) {
        uint32_t o[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) o[u] = 0x3c003c00u + u + t;
        if (mode == 3) {
          uint32_t r[16];
          ldg256(res + (size_t)m * N + n0 + j * 32, r);
          ldg256(res + (size_t)m * N + n0 + j * 32 + 16, r + 8);
#pragma unroll
          for (int u = 0; u < 16; ++u) o[u] += r[u];
        }
        stg256(out + (size_t)m * N + n0 + j * 32, o);
        stg256(out + (size_t)m * N + n0 + j * 32 + 16, o + 8);
      }
    } else if (mode == 1) {
      for (int j = h; j < BN / 32; j += 2) {
#pragma unroll
        for (int k8 = 0; k8 < 4; ++k8) {                       // 4 instructions x 8 rows
          const int m = m0 + k8 * 8 + (lane >> 2);
          uint4 v = make_uint4(0x3c003c00u + t, 1, 2, 3);
          *reinterpret_cast<uint4*>(out + (size_t)m * N + n0 + j * 32 + (lane & 3) * 8) = v;
        }
      }
    } else {                                                   // 64-column steps: columns [0,64) [64,128) then 32 left
      for (int j = h; j < 3; j += 2) {
        const int cols = (j < 2) ? 64 : 32;
        const int lpr = cols / 8;                              // lanes per row
        const int rpi = 32 / lpr;                              // rows per instruction
        for (int k8 = 0; k8 < 32 / rpi; ++k8) {
          const int m = m0 + k8 * rpi + lane / lpr;
          uint4 v = make_uint4(0x3c003c00u + t, 1, 2, 3);
          if (mode == 4) {
            const uint4 r = *reinterpret_cast<const uint4*>(res + (size_t)m * N + n0 + j * 64 + (lane % lpr) * 8);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
          }
          *reinterpret_cast<uint4*>(out + (size_t)m * N + n0 + j * 64 + (lane % lpr) * 8) = v;
        }
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0 + acc;
}
int main() {
  const int M = 16384 * 4;          // 512 m-tiles x 2 n-tiles = 1024 tiles
  __half *out, *res; long long* cyc;
  cudaMalloc(&out, (size_t)M * N * 2); cudaMalloc(&res, (size_t)M * N * 2); cudaMalloc(&cyc, 8);
  cudaMemset(res, 0, (size_t)M * N * 2);
  const int ctas = 128, tpc = 8;    // 1024 tiles
  for (int mode = 0; mode < 5; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) k<<<ctas, 256>>>(out, res, tpc, mode, cyc);
    cudaEventRecord(e0);
    for (int w = 0; w < 10; ++w) k<<<ctas, 256>>>(out, res, tpc, mode, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("mode %d: %.2f us per launch, %.0f ns per tile per CTA (%lld cycles/tile), %.0f GB/s written\n", mode, ms * 100,
           ms * 1e5 / tpc, c / tpc, (double)M * N * 2 / (ms / 10 * 1e-3) / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
