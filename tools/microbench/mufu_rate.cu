// MUFU / packed-FMA throughput on one SM (8 warps = 2 per SMSP, like the attention softmax warpgroups).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 mufu_rate.cu -o mufu_rate
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#define ITERS 4096
__global__ void k_ex2_f32(float* out, float x0, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = x0 + i * 0.01f + threadIdx.x * 1e-4f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ex2_f16x2(float* out, float x0, long long* cyc) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) { __half2 h = __floats2half2_rn(x0 + i * 0.01f, x0 - i * 0.01f); a[i] = *reinterpret_cast<uint32_t*>(&h); }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_ex2_f16(float* out, float x0, long long* cyc) {
  unsigned short a[8];
  for (int i = 0; i < 8; ++i) { __half h = __float2half_rn(x0 + i * 0.01f); a[i] = *reinterpret_cast<unsigned short*>(&h); }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16 %0, %0;" : "+h"(a[i]));
  }
  long long t1 = clock64();
  unsigned s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void k_ffma2(float* out, float x0, long long* cyc) {
  unsigned long long a[8], b, c;
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(0.999f), "f"(1.001f));
  asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(0.001f), "f"(-0.001f));
  for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(a[i]) : "f"(x0 + i), "f"(x0 - i));
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int i = 0; i < 8; ++i) s ^= a[i];
  out[threadIdx.x] = __uint_as_float((uint32_t)s);
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
}
__global__ void k_ffma(float* out, float x0, long long* cyc) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = x0 + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(0.001f));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
}
// mixed: 8 MUFU f32 + 24 FFMA2 per iteration (can the FMA pipe run under a saturated MUFU pipe?)
__global__ void k_mix(float* out, float x0, long long* cyc) {
  float a[8]; unsigned long long q[8], b, c;
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(0.999f), "f"(1.001f));
  asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(0.001f), "f"(-0.001f));
  for (int i = 0; i < 8; ++i) { a[i] = x0 + i * 0.01f; asm("mov.b64 %0, {%1,%2};" : "=l"(q[i]) : "f"(x0 + i), "f"(x0 - i)); }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(b), "l"(c));
      asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(q[(i + 3) & 7]) : "l"(b), "l"(c));
      asm volatile("fma.rn.ftz.f32x2 %0, %0, %1, %2;" : "+l"(q[(i + 5) & 7]) : "l"(b), "l"(c));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)q[i]);
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
}
// fp32 pair -> packed fp16x2 (F2FP.F16.F32.PACK_AB): which pipe, what rate?  (the softmax packs every probability)
__global__ void k_f2fp(float* out, float x0, long long* cyc) {
  float a[8]; uint32_t r[8];
  for (int i = 0; i < 8; ++i) { a[i] = x0 + i * 0.01f + threadIdx.x * 1e-4f; r[i] = 0; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t t;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      r[i] ^= t;
    }
  }
  long long t1 = clock64();
  uint32_t s = 0; for (int i = 0; i < 8; ++i) s ^= r[i];
  out[threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
}
// 6 MUFU + 4 F2FP per group of 8 elements, the softmax's mix: do they share a pipe?
__global__ void k_mufu_f2fp(float* out, float x0, long long* cyc) {
  float a[8]; uint32_t r[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) a[i] = x0 + i * 0.01f + threadIdx.x * 1e-4f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 6; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t t;
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a[6]), "f"(a[7]));
      r[i] ^= t;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[threadIdx.x] = s + __uint_as_float(r[0] ^ r[1] ^ r[2] ^ r[3]);
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMemset(cyc, 0, 64);
  for (int nthreads : {128, 256}) {
    k_ex2_f32<<<1, nthreads>>>(out, -0.5f, cyc); k_ex2_f16x2<<<1, nthreads>>>(out, -0.5f, cyc);
    k_ex2_f16<<<1, nthreads>>>(out, -0.5f, cyc); k_ffma2<<<1, nthreads>>>(out, 0.5f, cyc);
    k_ffma<<<1, nthreads>>>(out, 0.5f, cyc); k_mix<<<1, nthreads>>>(out, -0.5f, cyc);
    k_f2fp<<<1, nthreads>>>(out, 0.5f, cyc); k_mufu_f2fp<<<1, nthreads>>>(out, -0.5f, cyc);
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* names[8] = {"ex2.f32", "ex2.f16x2", "ex2.f16", "ffma2", "ffma", "mix(1 ex2 + 3 ffma2)", "f2fp (cvt f16x2)",
                            "group of 6 ex2 + 4 f2fp (per-group cycles x 8)"};
    const double warps_per_smsp = nthreads / 128.0;
    for (int i = 0; i < 8; ++i)
      printf("threads=%d %-22s %8lld cycles  %.2f cycles per warp-instr-group per SMSP (8*ITERS groups x %g warps)\n", nthreads,
             names[i], h[i], (double)h[i] / (8.0 * ITERS * warps_per_smsp), warps_per_smsp);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
