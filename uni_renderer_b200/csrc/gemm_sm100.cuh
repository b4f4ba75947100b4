// Implicit-GEMM convolution / linear kernel family for sm_100a (tcgen05 + TMEM + TMA).
//
//   out[m, n] = epilogue( sum_k A[m, k] * Wt[n, k] )
//
// A is an NHWC fp16 activation (or a plain row-major [M, K] matrix); the K loop walks a short list of "segments",
// each a (tensor map, tap pattern, channel blocks) triple, so that 3x3 / 1x1 / stride-2 convolutions, virtual
// channel concatenation (two sources) and a 1x1 shortcut fused behind a 3x3 conv all run as ONE accumulation in
// TMEM with no im2col buffer: a tap is just the same TMA box shifted by (dy, dx) with out-of-bounds zero fill.
// Wt is the packed weight matrix [N, Ktot] (K-major) whose K order equals the segment/tap/channel-block walk.
#pragma once
#include "common.cuh"

namespace unib {

constexpr int kBM = 128;        // rows (pixels) per tile == UMMA M
constexpr int kBK = 64;         // fp16 elements per K block == one 128-byte swizzle row
constexpr int kMaxSeg = 4;
constexpr int kMaxAMaps = 6;

enum SegKind : int { SEG_1x1 = 0, SEG_3x3 = 1, SEG_3x3_S2 = 2, SEG_3x3_S2P0 = 3, SEG_UP2x2 = 4 };

struct ConvSeg {
  int tmap;   // first A tensor map of this segment (SEG_3x3_S2 / _S2P0 use 4 consecutive parity maps)
  int kind;   // SegKind
  int nkb;    // channel blocks of 64 per tap
  int ntaps;  // 1 or 9
};

enum EpiFlags : int {
  EPI_GEGLU = 1,       // tile columns [0,BN/2) * gelu(columns [BN/2,BN)) -> N/2 outputs
  EPI_OUT_NCHW = 2,    // scalar stores to out[b][n][hw] (fp32 if EPI_OUT_F32 else fp16); small N only
  EPI_OUT_F32 = 4,
  EPI_SILU = 8,
  EPI_AXPBY = 16,      // x_prev = c_out * acc' + c_x * aux   (scheduler update fused behind conv_out)
};

struct GemmParams {
  int M, N;                  // GEMM rows (B*H*W or tokens), output channels (before GEGLU halving)
  int W, H;                  // spatial dims used to decode a tile origin (linear: W = 1<<30, H = 1)
  int nseg;
  ConvSeg seg[kMaxSeg];
  int total_kb;              // sum over segments of ntaps * nkb
  int splits;                // split-K factor (>1 => fp32 partials, finalize kernel applies the epilogue)
  unsigned long long mul_tiles, mul_ntiles;   // floor(2^40 / d) + 1 for d = m_tiles * n_tiles, n_tiles (fast_div)
  int w_shift, h_shift;      // log2 of W, H
  int rpb_shift;             // log2(rows_per_batch) if that is a power of two, else -1
  int m_tiles, n_tiles;      // m_tiles counts 128-row tiles (cg = 1) or 256-row PAIR tiles (cg = 2)
  int cg;                    // CTAs per MMA: 2 = cta_group::2 CTA pairs (cluster of 2 along M)
  // epilogue
  const float* bias;         // [bias_rows][N] fp32 (row b used for rows of batch b when bias_bstride != 0)
  int bias_bstride;
  const int* bias_step;      // optional: bias table of step s = bias + s * bias_step_stride (device-side step counter)
  long long bias_step_stride;
  int rows_per_batch;        // H*W of the OUTPUT (for batch index of a row)
  const __half* res;         // optional residual [M, ldr]
  int ldr;
  void* out;                 // fp16 [M, ldc]   (or NCHW when EPI_OUT_NCHW)
  int ldc;
  float* partial;            // split-K workspace [splits][M][N] fp32
  int flags;
  const float* axpby;        // EPI_AXPBY: [steps][2] (c_out, c_x); row *axpby_step (row 0 if null)
  const int* axpby_step;
  const float* aux;          // EPI_AXPBY: x_t, NCHW fp32
  float* aux_out;            // EPI_AXPBY: x_{t-1}, NCHW fp32 (may alias aux)
  int axpby_n0;              // channels < axpby_n0 keep aux unchanged
  int mode;                  // GemmMode (gemm_sm100.cu): 0 vector NHWC fp16 epilogue, 1 GEGLU, 2 direct stores
  float* rowstats_out;       // producer of a LayerNorm input: [M][2 * n_tiles][2] (sum, sumsq) of the output rows
  const float* ln_rowstats;  // consumer of LayerNorm(x): [M][ln_parts][2]; epilogue applies rstd * (acc - mean * wsum)
  const float* ln_wsum;      // [N]
  int ln_parts;
  float ln_eps, ln_inv_c;
  // nearest-2x upsample folded into the following 3x3 conv (SEG_UP2x2): the N dimension is [4 output parities][Cout],
  // a tile's parity is nt >> up_shift; rows are LOW-resolution pixels and the epilogue scatters them to the
  // high-resolution image.  up_shift < 0: off.
  int up_shift;
  int up_cout;
  // G = 2 grouped launch (the two directions of the dual-stream residual exchange at one skip site): m-tiles
  // [mt_single, 2 * mt_single) belong to a second problem of identical shape with its own A / B tensor maps
  // (maps.a[1], maps.b2) and its own epilogue pointers.  dual = 0: off.
  int dual, mt_single;
  const float* bias2;
  const __half* res2;
  void* out2;
  float* gn_part2;
  float* gn_part;            // GroupNorm statistics of the output: [M / gn_rows][N / gn_gran][2] (sum, sumsq), or null
  int gn_gran;               // channels per micro-group (even, divides the N tile and N)
  int gn_rows;               // rows per partial block: 32, 64 or 128 (divides the rows of one sample)
  int pdl_early;             // 1: fire the PDL trigger right after CTA setup instead of at the end (unib200_set_pdl(2))
  long long* trace;          // debug (unib200_debug_set_trace): [0] = launch counter, then 16 stamps per launch
};

struct alignas(64) GemmMaps {
  CUtensorMap a[kMaxAMaps];
  CUtensorMap b;
  CUtensorMap b2;            // weights of the second problem of a dual launch
};

// host-side launcher (gemm_sm100.cu)
cudaError_t launch_gemm(const GemmMaps& maps, const GemmParams& p, int bn, int num_sms, cudaStream_t stream);
int gemm_pick_bn(int N, int flags);
bool gemm_pair_supported(int bn);
size_t gemm_smem_bytes(int bn);

}  // namespace unib
