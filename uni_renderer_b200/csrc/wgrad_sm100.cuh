// Training-side kernels (first slice): weight gradient of the implicit-GEMM conv / linear family,
// bias gradient, GroupNorm(+SiLU) backward.  See wgrad_sm100.cu / elementwise.cu.
#pragma once
#include "common.cuh"

namespace unib {

struct WgradParams {
  int M, N, C;               // pixels (B*H*W or rows), output channels, input channels
  int taps;                  // 9 (3x3, pad 1, stride 1) or 1 (1x1 / linear)
  int w_shift, h_shift;      // log2 of the image dims (linear: 30, 0)
  int m_blocks;              // ceil(M / 128)
  int n_tiles, c_tiles;      // ceil(N / 128), ceil(C / 128)
  int splits;                // pixel splits (fp32 partial slabs + fixed-order reduce when > 1)
  float* dw;                 // fp32 [N][taps][C]
  float* partial;            // fp32 [splits][N][taps][C]
};

struct alignas(64) WgradMaps {
  CUtensorMap dy;            // 4-D {N, W, H, B}, box {64, bw, bh, bb}
  CUtensorMap x;             // 4-D {C, W, H, B}, same box (shifted by the tap at load time, zero fill outside)
};

cudaError_t launch_wgrad(const WgradMaps& maps, const WgradParams& p, cudaStream_t stream);
cudaError_t launch_colsum(const __half* dy, int ld, int M, int N, float* db, cudaStream_t stream);
int wgrad_cin_tile();

struct GnBwdParams {
  const __half* x; int ldx;      // forward input [B*HW, C]
  const __half* dz; int ldz;     // gradient w.r.t. the (SiLU'd) output [B*HW, C]
  __half* dx; int lddx;          // gradient w.r.t. x
  const float* gamma; const float* beta;
  float* dgamma_part;            // [B][C] per-sample partials (summed over B by the caller, fixed order)
  float* dbeta_part;             // [B][C]
  int HW, C, G, silu;
  float eps;
};
cudaError_t launch_gn_backward(const GnBwdParams& p, int B, cudaStream_t stream);
cudaError_t launch_pack_master(const float* w, __half* out, int O, int I, int taps, int dgrad, cudaStream_t stream);
cudaError_t launch_wgrad_scatter_add(const float* dw, float* grad, int N, int taps, int C, cudaStream_t stream);
cudaError_t launch_sum_slabs(const float* partial, float* out, long long n, int slabs, cudaStream_t stream);
// dgamma must hold 2 * C floats: on return [dgamma | dbeta]; slabs: fp32 [max_slabs][2][C] scratch
cudaError_t launch_layernorm_backward(const __half* x, const __half* dy, __half* dx, const float* gamma, float* dgamma,
                                      float* dbeta, float* slabs, int max_slabs, int rows, int C, float eps,
                                      cudaStream_t stream);
cudaError_t launch_geglu(const __half* proj, const __half* dout, __half* out, long long rows, int inner, int backward,
                         cudaStream_t stream);
cudaError_t launch_softmax_backward(const __half* P, __half* dP, int rows, int n, int ld, float scale, cudaStream_t stream);
cudaError_t launch_cvt_f32_f16(const float* src, __half* dst, long long rows, int cols, int ld, cudaStream_t stream);
cudaError_t launch_silu_f16(const __half* x, const __half* dy, __half* out, long long n, cudaStream_t stream);
cudaError_t launch_pool2x2_sum(const __half* src, __half* dst, int B, int H, int W, int C, cudaStream_t stream);
cudaError_t launch_scatter2x(const __half* src, __half* dst, int B, int H, int W, int C, cudaStream_t stream);
cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                         float wd, int step, float grad_scale, cudaStream_t stream);

}  // namespace unib
