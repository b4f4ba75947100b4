/* unib200 -- C ABI of the B200-native (sm_100a) kernels behind the Uni-Renderer dual-stream denoising hot path.
 *
 * The reference has NO native/FFI boundary on this path: the boundary a drop-in must honour is the Python
 * nn.Module surface of models/controlnet.py (UNet2DConditionModel.forward :781, AttributeEncoderModel.forward :1657,
 * AttributeDecoderModel.forward :2342) whose arithmetic is executed by torch (cuDNN conv, cuBLAS GEMM, SDPA,
 * native GroupNorm/LayerNorm).  This header is therefore the boundary of OUR replacement for those torch leaf
 * calls: uni_renderer_b200/ (Python, mirrors the reference classes) binds these symbols with ctypes.
 * Each entry point names the reference call site(s) whose device work it replaces.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer unless stated;
 * activations are NHWC fp16 ([B*H*W, C] row-major); all calls are asynchronous on `stream` (a cudaStream_t);
 * return 0 on success, negative on error with unib200_last_error() giving the reason; never throws.
 * If `prog` is non-NULL the op is RECORDED into the program (tensor maps pre-encoded) instead of launched;
 * unib200_program_run / unib200_program_graph_launch replay the recorded list.
 */
#ifndef UNIB200_H_
#define UNIB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIB200_VERSION 100

typedef struct unib200_program unib200_program;

/* ---- library / device ------------------------------------------------------------------------------------- */
int unib200_version(void);
const char* unib200_last_error(void);                 /* thread-local, valid until the next failing call */
int unib200_device_info(int* num_sms, int* cc_major, int* cc_minor);

/* Programmatic dependent launch (OFF by default: measured as no gain inside the two-lane step graph):
 * when enabled (1 = trigger at CTA end, 2 = early trigger) every kernel is launched with the
 * programmatic-stream-serialization attribute and waits (griddepcontrol.wait) before touching global memory, so the
 * prologue of kernel N+1 overlaps the tail of kernel N inside a stream / CUDA-graph branch.  Applies to launches made
 * after the call (A/B runs; env UNIB200_PDL in the Python binding). */
void unib200_set_pdl(int enabled);

/* ---- programs (recorded op lists; optionally instantiated as a CUDA graph) --------------------------------- */
unib200_program* unib200_program_create(void);
void unib200_program_destroy(unib200_program* prog);
int unib200_program_num_launches(const unib200_program* prog);   /* kernels launched by one run */
int unib200_program_run(unib200_program* prog, void* stream);
/* Lanes: ops recorded after set_lane(l) run on lane l (0 = the caller's stream, 1..3 = side streams owned by the
 * program).  Lanes run concurrently between barriers; a barrier joins every lane into lane 0 and forks again, and the
 * end of the program is an implicit join.  The RGB stream and the attribute stream of the dual-stream step are
 * data-independent between exchanges (models/controlnet.py:1078-1087 vs :2446-2461), so they are recorded on two
 * lanes and become two parallel branches of the step's CUDA graph. */
int unib200_program_set_lane(unib200_program* prog, int lane);
int unib200_program_barrier(unib200_program* prog);
int unib200_program_graph_instantiate(unib200_program* prog, void* stream);   /* capture run() into a CUDA graph */
int unib200_program_graph_launch(unib200_program* prog, void* stream);
/* accounting + measurement of a recorded program: per-op kind, algorithmic FLOPs (2*MAC, unpadded) and HBM bytes
 * (operands read once, result written once); _profile replays the program `iters` times with a CUDA-event pair
 * around every op on `stream` and returns each op's mean device time in ms (ms_out[num_ops]); host-synchronous. */
enum { UNIB200_OP_OTHER = 0, UNIB200_OP_GEMM = 1, UNIB200_OP_ATTENTION = 2, UNIB200_OP_GROUPNORM = 3,
       UNIB200_OP_LAYERNORM = 4 };
int unib200_program_num_ops(const unib200_program* prog);
int unib200_program_op_info(const unib200_program* prog, int i, int* kind, double* flops, double* bytes,
                            int* launches);
const char* unib200_program_op_desc(const unib200_program* prog, int i);    /* shape summary, owned by the program */
int unib200_program_profile(unib200_program* prog, void* stream, int iters, float* ms_out);

/* ---- implicit-GEMM convolution / linear (tcgen05) ----------------------------------------------------------
 * Replaces: nn.Conv2d 3x3/1x1 inside ResnetBlock2D, Transformer2DModel.proj_in/out, Down/Upsample2D, conv_in,
 * conv_out, the 26 exchange zero-convs (models/controlnet.py:1019,1157,1754-1775,2456,2476,2521) and every
 * nn.Linear of Attention / FeedForward (diffusers leaves called from models/unet_2d_blocks.py:803,1207,2576).
 *   out[m, n] = epilogue( sum over segments/taps/channels  A_seg[pixel(m)+tap, c] * weight[n, k] )
 */
/* _S2: 3x3 stride 2 pad 1 (Downsample2D of the UNets); _S2P0: 3x3 stride 2 with the input padded by one pixel at the
 * bottom / right only (diffusers Downsample2D(padding=0) = F.pad(x, (0,1,0,1)) + conv, the AutoencoderKL encoder). */
enum { UNIB200_SEG_1x1 = 0, UNIB200_SEG_3x3 = 1, UNIB200_SEG_3x3_S2 = 2, UNIB200_SEG_3x3_S2P0 = 3,
       UNIB200_SEG_UP2x2 = 4 /* nearest-2x upsample folded into the following 3x3 conv, see unib200_gemm_desc */ };
enum {
  UNIB200_EPI_GEGLU = 1,     /* weight rows interleaved per N-tile: out = (a+ba) * gelu(g+bg); N_out = N/2        */
  UNIB200_EPI_OUT_NCHW = 2,  /* store NCHW (fp16, or fp32 with OUT_F32) instead of NHWC fp16                      */
  UNIB200_EPI_OUT_F32 = 4,
  UNIB200_EPI_SILU = 8,
  UNIB200_EPI_AXPBY = 16     /* scheduler update fused behind conv_out: see unib200_gemm_desc.axpby              */
};

typedef struct {
  const void* ptr;   /* NHWC fp16 base of this source                                                            */
  int C;             /* channels taken from this source                                                           */
  int ld;            /* elements between consecutive pixels (>= C; lets a source be a column slice)               */
  int kind;          /* UNIB200_SEG_*                                                                             */
} unib200_seg;

typedef struct {
  int M, N;                 /* output rows (B*H*W or tokens) and output channels (pre-GEGLU)                     */
  int B, H, W;              /* OUTPUT image dims for conv segments; H = W = 0 => plain [M, K] row-major A        */
  int nseg;
  unib200_seg seg[4];       /* accumulated in order; weight K layout = [seg][tap][ceil(C/64)*64]                  */
  const void* weight;       /* packed fp16 [N, Ktot]                                                              */
  const float* bias;        /* fp32 [N] or [B][N] (bias_bstride = N), may be NULL                                 */
  int bias_bstride;
  const int* bias_step;     /* optional device counter: the bias table of step s starts at bias + s * bias_step_stride */
  int64_t bias_step_stride; /* (time-embedding projections tabulated for every step of a sampling loop)             */
  const void* res;          /* optional fp16 residual [M, ldr] added before the activation                        */
  int ldr;
  void* out;                /* fp16 [M, ldc] (or NCHW, see flags)                                                 */
  int ldc;
  int flags;                /* UNIB200_EPI_*                                                                       */
  int splits;               /* split-K factor; 0 = choose automatically                                           */
  float* partial;           /* split-K fp32 workspace (may be NULL => no split-K)                                 */
  size_t partial_bytes;
  const float* axpby;       /* EPI_AXPBY: device [steps][2] (c_out, c_x) table, row = *axpby_step (or row 0)       */
  const int* axpby_step;
  const float* aux;         /* EPI_AXPBY: current latent x_t, NCHW fp32                                           */
  float* aux_out;           /* EPI_AXPBY: x_{t-1} NCHW fp32 (may alias aux)                                       */
  int axpby_first_channel;  /* channels below this keep aux unchanged (the clean mask group, pipeline.py:2691)    */
                            /* EPI_AXPBY with out != NULL: out receives, NHWC fp16 [M, ldc], the raw prediction   */
                            /* with the clean channels passed through (= cat(latents_mask, mask_pred), the        */
                            /* attribute input of the cycle pass, train/train.py:1393)                            */
  /* LayerNorm folded into the GEMMs around it (BasicTransformerBlock.norm1/2/3): the GEMM that PRODUCES the
   * normalised tensor writes per-row partial statistics, the GEMM that CONSUMES LayerNorm(x) takes raw x with
   * gamma folded into its weights and corrects in the epilogue:
   *   out[m,n] = rstd[m] * (acc[m,n] - mean[m] * ln_wsum[n]) + bias[n],   bias[n] = b[n] + sum_k beta[k] W[n,k]
   * -- no LayerNorm kernel, no normalised tensor in HBM.                                                          */
  float* rowstats_out;      /* producer: fp32 [M][2*ceil(N/BN)][2] (sum, sum of squares) of the stored rows, or NULL */
  const float* ln_rowstats; /* consumer: the producer's table, [M][ln_parts][2], or NULL                           */
  int ln_parts;
  const float* ln_wsum;     /* consumer: fp32 [N], sum_k of the (gamma-folded, fp16-rounded) weight row            */
  float ln_eps;
  int ln_C;                 /* channels the statistics were taken over                                             */
  /* UNIB200_SEG_UP2x2 (single segment): Upsample2D = F.interpolate(scale 2, "nearest") + conv3x3 (models/
   * unet_2d_blocks.py:2588,2701) as ONE GEMM over the LOW-resolution input: output pixel (2h+py, 2w+px) only sees the
   * 2 x 2 input pixels (h-1+py+ty, w-1+px+tx), so each of the 4 output parities is a 2x2 conv whose weights are sums
   * of the 3x3 taps (packed on the host: N = 4 * Cout rows ordered [parity][Cout], K = [4 taps][ceil(C/64)*64]).
   * M, B, H, W describe the INPUT; out is the fp16 [B * 2H * 2W, ldc] high-resolution image; bias has Cout entries.
   * 2.25x fewer MACs than convolving the materialised upsampled tensor, which is never written.  Needs the vector
   * epilogue and Cout a power-of-two multiple of the N tile; no residual, no split-K.                                 */
  /* GroupNorm statistics fused into the epilogue of the GEMM that PRODUCES a GroupNorm input (ResnetBlock2D norm1 /
   * norm2, Transformer2DModel.norm, conv_norm_out): per (block of gn_rows rows, micro-group of gn_gran channels) the
   * sum and the sum of squares of the output; unib200_groupnorm with part1 / part2 set consumes them, so a GroupNorm
   * is ONE launch that reads its input once.  Needs the vector epilogue (no split-K), N % gn_gran == 0, gn_gran even
   * and dividing the N tile, gn_rows in {32, 64, 128} dividing the rows of one sample.                              */
  float* gn_part;           /* fp32 [M / gn_rows][N / gn_gran][2], or NULL                                         */
  int gn_gran;
  int gn_rows;
} unib200_gemm_desc;

int unib200_conv_gemm(unib200_program* prog, const unib200_gemm_desc* desc, void* stream);
/* G = 2 grouped launch: two GEMMs of identical shape (same M, N, one 1x1 / linear segment of the same C; plain bias /
 * residual / GroupNorm-statistics epilogue) as ONE kernel whose tile walk covers both.  Used for the dual-stream residual
 * exchange: at every skip site the RGB stream needs skip_U + zc_enc(skip_A) (models/controlnet.py:1078-1087,1115) and the
 * attribute stream skip_A + zc_dec(skip_U) (:2446-2461,2476-2477) -- both directions of a site run in one kernel. */
int unib200_conv_gemm_dual(unib200_program* prog, const unib200_gemm_desc* d0, const unib200_gemm_desc* d1, void* stream);
size_t unib200_packed_k(int nseg, const unib200_seg* seg);      /* Ktot of the packed weight matrix               */
int unib200_pick_bn(int N, int flags);   /* N-tile width the kernel will use (EPI_GEGLU weights are interleaved per tile) */

/* ---- fused attention (tcgen05 flash attention, no mask) ----------------------------------------------------
 * Replaces F.scaled_dot_product_attention inside diffusers Attention (AttnProcessor2_0) for attn1/attn2 of every
 * BasicTransformerBlock.  q/k/v are fp16 row-major token matrices; head h uses columns [h*d, (h+1)*d).
 */
typedef struct {
  const void* q; int ldq;
  const void* k; int ldk;
  const void* v; int ldv;
  void* out; int ldo;
  int B, heads, Nq, Nk, d;
  float scale;
  float* lse2;               /* optional fp32 [B, heads, Nq]: log2 of every row's sum of exp2(scale * log2e * s) -- what
                                unib200_attention_backward needs from the forward (training); NULL in inference    */
} unib200_attn_desc;
int unib200_attention(unib200_program* prog, const unib200_attn_desc* desc, void* stream);
/* debug only: attention kernels launched after this call write clock64 stamps of CTA (0,0,0) into dev_buf
 * (>= 4*16*8 int64); NULL switches tracing off. */
void unib200_debug_set_trace(void* dev_buf);

/* ---- GroupNorm(+SiLU) over NHWC with optional second source (virtual torch.cat, unet_2d_blocks.py:2546,2677) - */
typedef struct {
  const void* x1; int ld1; int C1;
  const void* x2; int ld2; int C2;
  int B, HW, groups;
  float eps;
  const float* gamma; const float* beta;
  void* out;                 /* fp16 [B*HW, C1+C2]                                                                */
  int silu;
  float* scratch;            /* fp32 scratch, scratch_floats >= B * chunks * groups * 2 for some chunks >= 1       */
  size_t scratch_floats;
  /* statistics already produced by the GEMM epilogues that wrote x1 / x2 (unib200_gemm_desc.gn_part): when part1 is
   * set (and part2 whenever x2 is), no statistics pass runs -- one launch, one read of the input                     */
  const float* part1;        /* [B * HW / part_rows][C1 / part_gran][2]                                            */
  const float* part2;        /* [B * HW / part_rows][C2 / part_gran][2]                                            */
  int part_gran, part_rows;
} unib200_gn_desc;
int unib200_groupnorm(unib200_program* prog, const unib200_gn_desc* desc, void* stream);

/* ---- LayerNorm over channels of [rows, C] fp16 (BasicTransformerBlock.norm1/2/3) ---------------------------- */
int unib200_layernorm(unib200_program* prog, const void* x, void* y, const float* gamma, const float* beta, int rows,
                      int C, float eps, void* stream);

/* ---- layout / resampling ------------------------------------------------------------------------------------ */
int unib200_to_nhwc(unib200_program* prog, const void* src, int src_is_f32, void* dst, int B, int C, int H, int W,
                    int64_t sb, int64_t sc, int64_t sh, int64_t sw, int Cpad, void* stream);
int unib200_from_nhwc(unib200_program* prog, const void* src, void* dst, int dst_is_f32, int B, int C, int HW, int ld,
                      void* stream);
int unib200_upsample2x(unib200_program* prog, const void* src, void* dst, int B, int H, int W, int C, void* stream);

/* ---- timestep embedding (Timesteps + TimestepEmbedding + all time_emb_proj, controlnet.py:909-916) ---------- */
int unib200_timestep_sinusoid(unib200_program* prog, const float* t, const int* step_idx, int t_stride, float* out,
                              int B, int dim, void* stream);
int unib200_gemv(unib200_program* prog, const float* x, const void* w_fp16, const float* bias, float* y, int B, int K,
                 int N, int act_silu, void* stream);

/* ---- scheduler update x_prev = c_out*model_out + c_x*x (DDIMScheduler.step, eta=0; pipeline.py:1649,2725) ---- */
int unib200_axpby(unib200_program* prog, const float* model_out, const float* x, float* out, const float* coef,
                  const int* step_idx, int64_t n, void* stream);
int unib200_add_int(unib200_program* prog, int* p, int v, void* stream);

/* ---- UniPCMultistepScheduler.step (order 2, bh2, predict-x0; what eval/test_real.py:485-493 attaches to every stream)
 * as ONE fused pass per stream and step: convert_model_output + corrector + history shift + predictor are linear
 * combinations with host-tabulated scalars, coef = device [steps][10] (scheduler.py UniPCSchedule), row *step_idx.
 * sample / last_sample / hist0 / hist1 / model_out: fp32 [B, C, HW]; channels < first_channel are left untouched. */
int unib200_unipc_step(unib200_program* prog, const float* model_out, float* sample, float* last_sample, float* hist0,
                       float* hist1, const float* coef, const int* step_idx, int B, int C, int HW, int first_channel,
                       void* stream);

/* ---- out = a + b over contiguous fp16 (skip + external residual when the three modules are called separately,
 *      models/controlnet.py:1084,1115; the fused step folds these adds into the zero-conv GEMM epilogue) -------- */
int unib200_add_f16(unib200_program* prog, const void* a, const void* b, void* out, int64_t n, void* stream);

/* ---- AutoencoderKL around the loop (models/pipeline.py:1531-1556 encode, :1664,2335-2344 decode)
 * The VAE's convolutions / GroupNorms / upsampling run on the entry points above.  Its one attention (mid block, a
 * single head of d = 512, diffusers Attention with residual_connection) does not fit the flash kernel's TMEM layout
 * and runs as S = Q K^T (unib200_conv_gemm with K as the [N, K] operand), this in-place row softmax
 * P = softmax(scale * S) over fp16 [rows, ld], and O = P V (unib200_conv_gemm with V^T as the [N, K] operand).
 * n is the number of real columns (any value, e.g. the 77 text tokens of the training path's cross-attention); ld is a
 * multiple of 8 covering n rounded up to 8, and the padding columns n .. round8(n) are written as zeros. */
int unib200_softmax_rows(unib200_program* prog, void* s_fp16, int rows, int n, int ld, float scale, void* stream);
/* DiagonalGaussianDistribution of `vae.encode(x).latent_dist`: moments fp32 [B, 2C, HW] (mean | logvar) ->
 * out fp32 [B, C, HW] = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scale; noise == NULL: mode() * scale. */
int unib200_gaussian_sample(unib200_program* prog, const float* moments, const float* noise, float* out, int B, int C,
                            int HW, float scale, void* stream);

/* ---- training, first slice (train/train.py:1324-1427 runs the three modules forward AND backward)
 * Data gradients of the conv / linear family need no kernel of their own: dX = conv(dY, W') is unib200_conv_gemm with the
 * weights repacked (taps flipped, in / out channels swapped).  New here: the WEIGHT gradient (pixels are the contraction
 * dimension: both operands are read MN-major straight from their NHWC tensors), the bias gradient, and the backward of
 * GroupNorm(+SiLU). */
typedef struct {
  const void* x; int C; int ldx;      /* forward input of the conv, fp16 NHWC [M, ldx] (first C channels)                 */
  const void* dy; int N; int lddy;    /* gradient w.r.t. the conv output, fp16 [M, lddy]                                  */
  int M, B, H, W;                     /* M = B*H*W; H = W = 0: plain [M, K] matrices (linear layer)                        */
  int taps;                           /* 9 = 3x3 pad 1 stride 1, 1 = 1x1 / linear                                          */
  float* dw;                          /* out: fp32 [N][taps][C] (the packed K order of the forward weights, unpadded)      */
  float* db;                          /* out (optional): fp32 [N] = column sums of dy                                      */
  float* partial; size_t partial_bytes;   /* fp32 workspace for pixel splits (may be NULL => one split)                    */
} unib200_wgrad_desc;
int unib200_conv_wgrad(unib200_program* prog, const unib200_wgrad_desc* desc, void* stream);
typedef struct {
  const void* x; int ldx;             /* forward input of the GroupNorm, fp16 [B*HW, ldx]                                  */
  const void* dz; int ldz;            /* gradient w.r.t. silu(groupnorm(x)) (or groupnorm(x) when silu = 0)                */
  void* dx; int lddx;                 /* out: gradient w.r.t. x, fp16                                                      */
  const float* gamma; const float* beta;
  float* dgamma; float* dbeta;        /* out: fp32 [C]                                                                     */
  float* scratch;                     /* fp32 [2 * B * C]                                                                  */
  int B, HW, C, groups, silu;
  float eps;
} unib200_gn_bwd_desc;
int unib200_groupnorm_backward(unib200_program* prog, const unib200_gn_bwd_desc* desc, void* stream);
/* out[n] = sum over the M rows of the fp16 matrix x [M, ld] (first N columns, N % 8 == 0), fp32, fixed order: bias
 * gradients, and the gradient of a per-sample broadcast add (the time-embedding add of ResnetBlock2D). */
int unib200_colsum(unib200_program* prog, const void* x, int ld, int M, int N, float* out, void* stream);
/* Training glue between the fp32 master weights and the kernels (the reference keeps fp32 parameters under
 * accelerate's fp16 autocast, train/train.py:1324-1427; torch's permute / flip / pad / cast chains cost ~25 ms per
 * step).  w: fp32 [O][I][taps] (taps 1 = Linear / 1x1, 9 = 3x3, the reference's Conv2d layout).
 * dgrad = 0: out = fp16 [O][taps * Ipad], the K-major operand unib200_conv_gemm takes (Ipad = I rounded up to 64);
 * dgrad = 1: out = fp16 [I][taps * Opad], the operand of the data-gradient convolution
 *            dX = conv(dY, W'), W'[i][o][ky][kx] = W[o][i][2 - ky][2 - kx].
 * Only the valid columns are written: the caller zeroes the buffer once when it allocates it. */
int unib200_pack_master_weight(unib200_program* prog, const float* w, int O, int I, int taps, void* out, int dgrad,
                               void* stream);
/* grad[n][c][t] += dw[n][t][c]: unib200_conv_wgrad's tap-major fp32 output added into a reference-layout gradient */
int unib200_wgrad_scatter_add(unib200_program* prog, const float* dw, float* grad, int N, int taps, int C, void* stream);

/* transformer-block backward helpers (BasicTransformerBlock: LayerNorm, GEGLU, softmax of a materialised attention) */
/* LayerNorm backward over [rows, C] fp16: dx, and dgamma_dbeta = fp32 [2 * C] = [dgamma | dbeta]; scratch fp32
 * [scratch_floats], at least 2 * C */
int unib200_layernorm_backward(unib200_program* prog, const void* x, const void* dy, void* dx, const float* gamma,
                               float* dgamma_dbeta, float* scratch, size_t scratch_floats, int rows, int C, float eps,
                               void* stream);
/* GEGLU over proj fp16 [rows, 2 * inner] = [a | g]: forward out = a * gelu(g) (dout = NULL), backward out = dproj given
 * dout [rows, inner] */
int unib200_geglu(unib200_program* prog, const void* proj, const void* dout, void* out, int64_t rows, int inner, void* stream);
/* dS = scale * P o (dP - rowsum(dP o P)), written over dP; fp16 [rows, ld], first n columns */
int unib200_softmax_backward(unib200_program* prog, const void* P, void* dP, int rows, int n, int ld, float scale, void* stream);
/* fp32 [rows, cols] contiguous -> fp16 with leading dimension ld */
int unib200_cvt_f32_f16(unib200_program* prog, const float* src, void* dst, int64_t rows, int cols, int ld, void* stream);
/* Flash-attention backward (head dims <= 80): gradients of O = softmax(Q K^T scale) V w.r.t. Q, K, V from dO without
 * materialising the Nq x Nk matrices (csrc/attention_bwd_sm100.cu).  q / k / v / o / dout are the forward's fp16
 * matrices (head h in columns [h*d, (h+1)*d)), lse2 the forward's optional output; D is fp32 scratch [B*heads*Nq];
 * dq_acc is an fp32 [B*Nq, ld_dq] accumulator that the CALLER zeroes (every key block adds into it with atomics:
 * dQ is not bit-reproducible run to run); dk / dv are written.  Replaces the autograd of
 * F.scaled_dot_product_attention under accelerator.backward (train/train.py:1421). */
typedef struct {
  const void* q; int ldq;
  const void* k; int ldk;
  const void* v; int ldv;
  const void* o; int ldo;
  const void* dout; int lddo;
  const float* lse2;
  float* D;
  float* dq_acc; int ld_dq;
  void* dk; int ld_dk;
  void* dv; int ld_dv;
  int B, heads, Nq, Nk, d;
  float scale;
} unib200_attn_bwd_desc;
int unib200_attention_backward(unib200_program* prog, const unib200_attn_bwd_desc* desc, void* stream);
/* network-level training glue (uni_renderer_b200/trainer.py; train/train.py:1324-1427):
 * SiLU on n fp16 elements (dy NULL: out = silu(x); else out = dy * silu'(x)) -- the time-embedding MLP's activations */
int unib200_silu_f16(unib200_program* prog, const void* x, const void* dy, void* out, int64_t n, void* stream);
/* adjoint of nearest-2x upsampling (F.interpolate backward, models/unet_2d_blocks.py:2588): dst [B,H,W,C] = 2x2 block
 * sums of src [B,2H,2W,C]; fp16 NHWC, C a multiple of 8 */
int unib200_pool2x2_sum(unib200_program* prog, const void* src, void* dst, int B, int H, int W, int C, void* stream);
/* zero insertion dst [B,2H,2W,C] <- src [B,H,W,C] at even pixels: the gradients of the stride-2 Downsample2D conv become
 * stride-1 problems (dX = conv3x3(scatter(dY), flipped W); dW = wgrad(x, scatter(dY))) */
int unib200_scatter2x(unib200_program* prog, const void* src, void* dst, int B, int H, int W, int C, void* stream);
/* AdamW on flat fp32 buffers with torch.optim.AdamW's arithmetic (train/train.py:1424 optimizer.step()); gradients are
 * multiplied by grad_scale first (inverse loss scale x clip coefficient); step counts from 1 */
int unib200_adamw_step(unib200_program* prog, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- step-level context -----------------------------------------------------------------
 * A context owns what one dual-stream sampler needs at run time -- recorded programs (ownership passes to it), device
 * buffers it allocated or was handed, weights uploaded into it -- so that, once a plan exists, a denoising step, a
 * per-network forward or the WHOLE sampling loop is one C call with plain pointers: no Python, no torch in the loop.
 * (Recording the programs = the layer wiring of models/controlnet.py + unet_2d_blocks.py stays in the host layer,
 * uni_renderer_b200/engine.py + pipeline.py, which calls the op entry points above.)
 * Replaces, at the reference's boundary: the three module forwards (models/controlnet.py:781,1657,2342) ->
 * unib200_ctx_run(ctx, "unet" | "attr_enc" | "attr_dec"); one loop body (models/pipeline_new_d4p.py:1391-1453,
 * models/pipeline.py:1586-1653, 2627-2733) -> unib200_dual_step; the whole `for t in timesteps` loop ->
 * unib200_sample_loop.  One thread at a time per context; asynchronous w.r.t. the host like every other entry point. */
typedef struct unib200_ctx unib200_ctx;
typedef struct {
  int use_graph;            /* replay the step program as a CUDA graph when it has been instantiated (default 1)       */
  int reserved[7];
} unib200_config;
enum { UNIB200_F16 = 0, UNIB200_F32 = 1, UNIB200_I32 = 2 };
unib200_ctx* unib200_create(int device, const unib200_config* cfg);            /* cfg may be NULL                      */
void unib200_destroy(unib200_ctx* ctx);                                        /* frees programs, buffers, weights     */
/* copy a tensor (host or device memory) into context-owned device memory under `key`; the caller keeps `src` */
int unib200_load_weight(unib200_ctx* ctx, const char* key, const void* src, int dtype, const int64_t* shape, int ndim);
int unib200_alloc(unib200_ctx* ctx, const char* key, size_t bytes, int zero);  /* named context-owned device buffer    */
int unib200_bind(unib200_ctx* ctx, const char* key, void* dev_ptr, size_t bytes);   /* borrow a caller-owned buffer   */
void* unib200_buffer(unib200_ctx* ctx, const char* key, size_t* bytes);        /* device pointer of `key`, or NULL      */
/* hand a recorded program to the context under a name; "once" (schedule tables, runs at attach time of a loop),
 * "setup" (step-invariant work, once per call) and "step" (one denoising step) drive the loop entry points */
int unib200_ctx_attach(unib200_ctx* ctx, const char* name, unib200_program* prog);
int unib200_ctx_run(unib200_ctx* ctx, const char* name, void* stream);
int unib200_unet_forward(unib200_ctx* ctx, void* stream);                      /* = unib200_ctx_run(ctx, "unet")       */
int unib200_attr_enc_forward(unib200_ctx* ctx, void* stream);
int unib200_attr_dec_forward(unib200_ctx* ctx, void* stream);
/* ONE denoising step of the attached "step" program (both streams, exchange, scheduler updates; the device-side step
 * counter advances) */
int unib200_dual_step(unib200_ctx* ctx, void* stream);
/* the whole loop: copies the inputs (host or device pointers; NULL = keep the buffer's content) into the buffers bound
 * as "lat_img" / "lat_attr" / "ehs", zeroes "step", runs "setup" and n_steps x "step", and copies the final latents out
 * (NULL = leave them in the bound buffers) -- everything enqueued on `stream`. */
int unib200_sample_loop(unib200_ctx* ctx, int n_steps, const void* lat_img, const void* lat_attr, const void* ehs,
                        void* out_lat_img, void* out_lat_attr, void* stream);
/* The sharded loop's one collective, ncclAllGather of the final latents, is issued by the host layer through
 * torch.distributed (pipeline.all_gather_latents): this library does not link NCCL. */

#ifdef __cplusplus
}
#endif
#endif /* UNIB200_H_ */
