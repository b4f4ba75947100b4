// HBM-bound helper kernels (see elementwise.cu).
#pragma once
#include "common.cuh"

namespace unib {

struct GnParams {
  const __half* x1; int ld1; int C1;   // first source  [B*HW, ld1], channels [0, C1)
  const __half* x2; int ld2; int C2;   // optional second source (virtual channel concat), channels [C1, C1+C2)
  int HW, G;
  float eps;
  const float* gamma; const float* beta;
  __half* out;                          // [B*HW, C1+C2]
  int silu;
  float* partial;                       // [B][max_chunks][G][2] fp32 scratch
  int max_chunks;
  int stat_chunks;                      // filled in by the launcher
  // statistics from the producing GEMMs' epilogues (GemmParams::gn_part): [B*HW / part_rows][C{1,2} / part_gran][2]
  const float* part1; const float* part2;
  int part_gran, part_rows;
  int part_split;            // launcher: threads sharing one statistics chain (sizes the kernel's scratch)
};

cudaError_t launch_groupnorm(const GnParams& p, int B, int num_sms, cudaStream_t stream);
cudaError_t launch_groupnorm_parts(const GnParams& p, int B, int num_sms, cudaStream_t stream);
cudaError_t launch_layernorm(const __half* x, __half* y, const float* gamma, const float* beta, int rows, int C,
                             float eps, cudaStream_t stream);
cudaError_t launch_to_nhwc(const void* src, int src_is_f32, __half* dst, int B, int C, int H, int W, long long sb,
                           long long sc, long long sh, long long sw, int Cpad, cudaStream_t stream);
cudaError_t launch_from_nhwc(const __half* src, void* dst, int dst_is_f32, int B, int C, int HW, int ld,
                             cudaStream_t stream);
cudaError_t launch_upsample2x(const __half* src, __half* dst, int B, int H, int W, int C, cudaStream_t stream);
cudaError_t launch_timestep_sinusoid(const float* t, const int* step_idx, int t_stride, float* out, int B, int dim,
                                     cudaStream_t stream);
cudaError_t launch_gemv(const float* x, const __half* Wt, const float* bias, float* y, int B, int K, int N,
                        int act_silu, cudaStream_t stream);
cudaError_t launch_axpby(const float* model_out, const float* x, float* out, const float* coef, const int* step_idx,
                         long long n, cudaStream_t stream);
cudaError_t launch_unipc_step(const float* out, float* S, float* LS, float* H0, float* H1, const float* coef,
                              const int* step_idx, int B, int C, int HW, int c_first, cudaStream_t stream);
cudaError_t launch_add_f16(const __half* a, const __half* b, __half* out, long long n, cudaStream_t stream);
cudaError_t launch_add_int(int* p, int v, cudaStream_t stream);
cudaError_t launch_softmax_rows(__half* S, int rows, int n, int ld, float scale, cudaStream_t stream);
cudaError_t launch_gaussian_sample(const float* moments, const float* noise, float* out, int B, int C, int HW,
                                   float scale, cudaStream_t stream);

}  // namespace unib
