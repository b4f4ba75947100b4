// FlashAttention-style fused softmax(Q K^T * scale) V for sm_100a (tcgen05 + TMEM + TMA), no mask.
//
// One CTA = TWO 128-row query tiles of one (batch, head) that ping-pong on the tensor core: while softmax warpgroup 0
// exponentiates S0(j), the MMA warp runs S1(j) = Q1 K(j)^T and O0 += P0(j-1) V(j-1), and vice versa.  S = Q K^T and
// O += P V both run on tcgen05 with the accumulators in TMEM (per tile: S BKV fp32 columns, O round16(d) columns).
// The online softmax runs one thread per query row (tcgen05.ld 32x32b gives each thread its own row, so no cross-lane
// reductions) in a SINGLE pass over S held in registers, with fp32 statistics; O is rescaled in TMEM only when a
// row's running maximum moved by more than 2^8 (lazy rescale; the final normalisation is exact either way).  P goes
// back to shared memory as fp16 in the 128B-swizzled K-major layout and feeds the second MMA.
// At head dim 40 the kernel is bound by the MUFU unit (ex2: 16/clk/SM on B200), not by the tensor core: the two
// warpgroups exist to keep MUFU busy while the other tile waits on its MMAs.
// Warp roles (320 threads): warps 0-3 softmax/epilogue of tile 0, warps 4-7 of tile 1, warp 8 TMA producer,
// warp 9 MMA issuer + TMEM owner.
// Q/K/V are read straight out of the (fused) projection outputs through 4-D tensor maps {d, tokens, heads, batch};
// head dims that are not multiples of 64 rely on TMA out-of-bounds zero fill, so nothing is padded in HBM.
#include "attention_sm100.cuh"

#include <stdlib.h>

// measured on B200 (tools/ab_attn.sh, 4096 x 4096 tokens, d = 40): 0 -> 241.9 us, 1 -> 251.6, 2 -> 248.6, 3 -> 241.1,
// 5 -> 230.5, 7 -> 219.5, 9 -> 249.4, 11 -> 245.1; at d = 80 (BKV = 64) variant 0 stays the fastest (25.6 vs 27.1 us).
// With P in tensor memory (bit 64; profiles/r2z_ab_attention_p_tmem.txt): 71 -> 195.4, 87 -> 201.9, 43091 -> 194.0,
// 43219 -> 192.8 (within run-to-run noise of 71, which needs neither the ones tile nor d <= 48)
#ifndef UNIB_ATTN_DEFAULT_VARIANT
#define UNIB_ATTN_DEFAULT_VARIANT 71
#endif

namespace unib {

template <int NCH>
struct AttnCfg {
  static constexpr int kBKV = (NCH == 1) ? 128 : 64;       // kv rows per block
  static constexpr int BKV_ones_bytes() { return kBKV * 128; }
  static constexpr int kQBytes = NCH * 128 * 128;          // per tile: NCH chunks of [128 rows x 64 fp16]
  static constexpr int kKBytes = NCH * kBKV * 128;         // NCH chunks of [BKV rows x 64 fp16]
  static constexpr int kStageBytes = 2 * kKBytes;          // K + V
  static constexpr int kPBytes = (kBKV / 64) * 128 * 128;  // per tile: [128 rows x BKV fp16] as 64-wide chunks
  static constexpr int kStages = (NCH == 3) ? 2 : 3;
  static constexpr int kKvOff = 2 * kQBytes;
  static constexpr int kPOff = kKvOff + kStages * kStageBytes;
  static constexpr int kBarOff = kPOff + 2 * kPBytes;
  static constexpr int kSmemBytes = kBarOff + 256;         // base is declared 1024-aligned (checked at run time)
  // tensor-core row sums (variant bit 16, NCH == 1): a [BKV x 64] fp16 tile of ONES, addressed like one V chunk
  static constexpr int kOnesOff = kBarOff + 1024;
  static constexpr int kSmemBytesOnes = kOnesOff + BKV_ones_bytes();
  static constexpr int kTmemCols = 512;                    // 2 x S (BKV) + 2 x O (<= 64 | 128 | 192) [+ 2 x 16 row sums]
  static_assert(kSmemBytes + 1024 <= 232448, "shared memory budget");
  static_assert(NCH != 1 || kSmemBytesOnes + 1024 <= 232448, "shared memory budget (ones tile)");
};

// Fully unrolled MMA issue sequences: descriptors differ from a precomputed base only by compile-time constants, so
// the issuing thread spends one add per operand between tcgen05.mma instructions (the issue loop, not the tensor
// pipe, was the limiter when descriptors were rebuilt per instruction).
template <int KS, int BKV>
__device__ __forceinline__ void issue_qk(uint32_t d_tmem, uint64_t q_desc, uint64_t k_desc, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int ch = ks >> 2, within = ks & 3;
    umma_f16_ss(d_tmem, q_desc + static_cast<uint64_t>((ch * 16384 + within * 32) >> 4),
                k_desc + static_cast<uint64_t>((ch * (BKV * 128) + within * 32) >> 4), idesc, ks > 0 ? 1u : 0u);
  }
}
template <int BKV>
__device__ __forceinline__ void issue_qk_dyn(int ks_count, uint32_t d_tmem, uint64_t q_desc, uint64_t k_desc,
                                             uint32_t idesc) {
  switch (ks_count) {
    case 1: issue_qk<1, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 2: issue_qk<2, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 3: issue_qk<3, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 4: issue_qk<4, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 5: issue_qk<5, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 6: issue_qk<6, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 7: issue_qk<7, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 8: issue_qk<8, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 9: issue_qk<9, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 10: issue_qk<10, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    case 11: issue_qk<11, BKV>(d_tmem, q_desc, k_desc, idesc); break;
    default: issue_qk<12, BKV>(d_tmem, q_desc, k_desc, idesc); break;
  }
}
template <int BKV>
__device__ __forceinline__ void issue_pv(uint32_t d_tmem, uint64_t p_desc, uint64_t v_desc, uint32_t idesc, bool first) {
#pragma unroll
  for (int ks = 0; ks < BKV / 16; ++ks) {
    const int ch = ks >> 2, within = ks & 3;
    umma_f16_ss(d_tmem, p_desc + static_cast<uint64_t>((ch * 16384 + within * 32) >> 4),
                v_desc + static_cast<uint64_t>((ks * 2048) >> 4), idesc, (!first || ks > 0) ? 1u : 0u);
  }
}

// O += P V with P read from TENSOR MEMORY (variant bit 64): K step ks reads the 8 columns [8 ks, 8 ks + 8) of P_t
template <int BKV>
__device__ __forceinline__ void issue_pv_ts(uint32_t d_tmem, uint32_t p_tmem, uint64_t v_desc, uint32_t idesc, bool first) {
#pragma unroll
  for (int ks = 0; ks < BKV / 16; ++ks)
    umma_f16_ts(d_tmem, p_tmem + ks * 8, v_desc + static_cast<uint64_t>((ks * 2048) >> 4), idesc, (!first || ks > 0) ? 1u : 0u);
}

// debug stamps: slot layout trace[(who * 16 + j) * 8 + k]
#ifdef UNIB_ATTN_TRACE
#define ATTN_TRACE(who, j, k)                                                                                   \
  do {                                                                                                          \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 16)                \
      p.trace[((who) * 16 + (j)) * 8 + (k)] = clock64();                                                        \
  } while (0)
#else
#define ATTN_TRACE(who, j, k) do { } while (0)
#endif

// VAR (softmax variant bits; A/B-measured on B200):
//   1  packed fp32 pairs: the scale/subtract runs as FFMA2, the row sums as FADD2, the row maximum as FMNMX3
//      (halves the issue slots of the non-MUFU work of the exponential loop)
//   2  one-time stagger: warpgroup 1 starts its first block when warpgroup 0 is half-way through its exponentials, so
//      the two tiles alternate on the MUFU unit instead of running their exponential phases in lockstep
//   4  every 4th pair of exponentials on the FMA pipe (exp2_poly_x2);  8  every 2nd pair
//  16  row sums on the TENSOR CORE: a third MMA per block, L_t += P_t(j) 1 with an all-ones [BKV x 16] B operand, keeps
//      the softmax denominators in 16 TMEM columns next to O_t (rescaled with it) -- the per-element FADD leaves the
//      exponential loop, and the denominator sums exactly the fp16 probabilities the P V product uses
//  32  early S: the tcgen05.ld of S_t(j+1) is issued before the P_t(j) stores / fence / arrive, so the TMEM load
//      latency hides behind them instead of heading the next block
//  bits 8..15: explicit polynomial pattern over the 8 pairs of two consecutive 8-element groups (bit (u & 1) * 4 + e);
//      bits 4 / 8 are the patterns 0x88 / 0xAA
//  64  P in TENSOR MEMORY (the FlashAttention-4 arrangement): the softmax threads write their fp16 probabilities with
//      tcgen05.st into 64 TMEM columns per tile and the P V product takes its A operand from there.  The shared-memory
//      P round trip (32 KB written + 32 KB read per tile and block, of ~116 KB in total) disappears: at head dim 40 the
//      kernel was bound by the 128 B/clk shared-memory port (1812 predicted vs 1830 measured cycles per block pair)
// 128  (with 64) the first half of a row's probabilities leaves for P_t half-way through the exponential loop, which
//      frees 32 registers for the second half (the 3-of-8 / every-2nd-pair polynomial variants spill without it)
constexpr int kAttnPacked = 1, kAttnStagger = 2, kAttnPoly4 = 4, kAttnPoly2 = 8, kAttnTcSum = 16, kAttnEarlyS = 32,
              kAttnPTmem = 64, kAttnHalfStore = 128;
__host__ __device__ constexpr int attn_poly_pattern(int var) {
  return ((var >> 8) & 0xFF) ? ((var >> 8) & 0xFF) : (var & kAttnPoly2) ? 0xAA : (var & kAttnPoly4) ? 0x88 : 0;
}
// zero-instruction register dependency: consumers of v[] cannot be scheduled above the tcgen05.wait::ld that precedes it
__device__ __forceinline__ void touch32(float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
  asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                    "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                    "+r"(r[30]), "+r"(r[31]));
}

// SPLIT = 2 (P in tensor memory, head dims <= 64): TWO warpgroups per query tile, each exponentiating one 64-column
// half of the tile's S block for the same 128 rows (warps w and w + 4 share a TMEM lane quadrant, like the GEMM
// epilogue): sixteen softmax warps = four per scheduler instead of two, against the per-warp dependency latency that
// bounds the exponential loop once the shared-memory port is out of the way.  The halves agree on the row maximum
// through shared memory and one named barrier per block and keep separate row sums.
// Register re-allocation between the warp roles (setmaxnreg, as in the FlashAttention-3/4 kernels): the register file
// is handed out per 4-warp group, so the 10-warp CTA is launched with 12 warps (two idle) at the 168 registers
// __launch_bounds__(384, 1) allows; the producer / MMA warpgroup then shrinks to 104 registers and each softmax
// warpgroup grows to 200 (the pool is the 3 x 168 of the launch: 2 x 208 + 96 deadlocks in TRY_ALLOC) -- S (128 fp32) and the packed P (64) of a row fit without spills and the exponential loop
// has room to interleave independent elements (a plain 320-thread launch cannot get more than 168: 184 and 192 fail
// with "too many resources requested for launch").  UNIB_ATTN_SETMAXNREG=0 builds the old 320-thread kernel.
// MEASURED SLOWER (profiles/r2z_ab_attention_p_tmem.txt: variant 71 195 -> 207 us, 67 196 -> 204, 43091 194 -> 196): with
// 200 registers ptxas schedules the exponential loop differently and loses more than the spills cost.  Off by default.
#ifndef UNIB_ATTN_SETMAXNREG
#define UNIB_ATTN_SETMAXNREG 0
#endif
template <int SPLIT>
constexpr int attn_threads() { return (SPLIT == 1 && UNIB_ATTN_SETMAXNREG) ? 384 : (8 * SPLIT + 2) * 32; }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

template <int NCH, int VAR, int SPLIT = 1>
__global__ void __launch_bounds__(attn_threads<SPLIT>(), 1)
attention_tcgen05_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
  using Cfg = AttnCfg<NCH>;
  constexpr int BKV = Cfg::kBKV;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) {                // SWIZZLE_128B tiles need a 1024 B aligned base (no static smem here)
    if (threadIdx.x == 0) printf("unib200: attention smem base 0x%x is not 1024-byte aligned\n", base);
    __trap();
  }
  const uint32_t kv_smem = base + Cfg::kKvOff;
  const uint32_t bar_base = base + Cfg::kBarOff;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto s_full = [&](int t) { return bar_base + 8u * (7 + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (9 + t); };
  auto o_ready = [&](int t) { return bar_base + 8u * (11 + t); };
  auto pv_done = [&](int t) { return bar_base + 8u * (13 + t); };
  auto s_free = [&](int t) { return bar_base + 8u * (15 + t); };     // softmax_t holds S_t(j) in registers
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kBarOff + 248);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int q_base = blockIdx.x * 256;
  const int ntile = (q_base + 128 < p.Nq) ? 2 : 1;     // second tile may not exist
  const int nblk = (p.Nk + BKV - 1) / BKV;
  const int dpad = (p.d + 15) & ~15;
  constexpr bool kTcSum = (VAR & kAttnTcSum) != 0, kEarlyS = (VAR & kAttnEarlyS) != 0;
  constexpr int kPat = attn_poly_pattern(VAR);
  constexpr bool kPTmem = (VAR & kAttnPTmem) != 0, kHalfStore = kPTmem && (VAR & kAttnHalfStore) != 0;
  // TMEM budget: NCH = 1: 2 x (S 128 + O 64 + P 64) = 512 columns; NCH = 2 (BKV = 64): 2 x (S 64 + O 128 + P 32 of 64)
  static_assert(!kPTmem || NCH <= 2, "P in tensor memory: head dims <= 128 only");
  static_assert(!kHalfStore || NCH == 1, "half-way P store: BKV = 128 only");
  static_assert(!kTcSum || NCH == 1, "tensor-core row sums: head dims <= 64 only (TMEM / shared-memory budget)");
  static_assert(!kEarlyS || (NCH == 1 && (VAR & kAttnPacked)), "early S: BKV = 128 and the packed row maximum");

  if (kTcSum) {                              // the ones tile (generic-proxy stores -> visible to the tensor core)
    uint4* ones = reinterpret_cast<uint4*>(smem + Cfg::kOnesOff);
    const uint4 one8 = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    for (int i = threadIdx.x; i < BKV * 128 / 16; i += blockDim.x) ones[i] = one8;
    fence_proxy_async_shared();
  }
  constexpr int kProdWarp = 8 * SPLIT, kMmaWarp = 8 * SPLIT + 1;
  static_assert(SPLIT == 1 || (SPLIT == 2 && NCH == 1 && (VAR & kAttnPTmem) && !(VAR & (kAttnTcSum | kAttnEarlyS | kAttnHalfStore))),
                "column-split softmax: P in tensor memory, head dims <= 64, plain variants only");
  if (warp == kProdWarp && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 128 * SPLIT);
      mbar_init(o_ready(s), 1);
      mbar_init(pv_done(s), 1);
      mbar_init(s_free(s), 128 * SPLIT);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.pdl_early) pdl_launch();
  pdl_wait();        // PDL: see common.cuh (trigger at the end of the CTA, as in the GEMM kernel)
  // TMEM columns: [S0 | S1 | O0 | O1]
  auto t_s_col = [&](int t) { return static_cast<uint32_t>(t * BKV); };
  // P in tensor memory AND row sums: 2 x (S 128 + O dpad + P 64 + L 16) columns only fit with dpad <= 48 (launcher)
  const int o_stride = (kPTmem && kTcSum) ? dpad : NCH * 64;
  auto t_o_col = [&](int t) { return static_cast<uint32_t>(2 * BKV + t * o_stride); };
  auto t_p_col = [&](int t) { return static_cast<uint32_t>(2 * BKV + 2 * o_stride + t * 64); };      // kPTmem: fp16 P_t
  auto t_l_col = [&](int t) { return static_cast<uint32_t>(2 * BKV + 2 * o_stride + (kPTmem ? 128 : 0) + t * 16); };

  constexpr bool kRegSplit = (SPLIT == 1 && UNIB_ATTN_SETMAXNREG);
  // the control warpgroup: producer, MMA issuer and (register split) the two idle warps that complete the warpgroup
  const bool ctrl_wg = kRegSplit ? (warp >= 8) : (warp == kProdWarp || warp == kMmaWarp);
  if (ctrl_wg) {
  if (kRegSplit) setmaxnreg_dec<104>();
  if (warp == kProdWarp) {
    // =============================== TMA producer ===============================
    // warp-uniform loop, one elected lane issues (operands stay in uniform registers)
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ntile * Cfg::kQBytes);
      for (int t = 0; t < ntile; ++t)
        for (int ch = 0; ch < NCH; ++ch)
          tma_load_4d(base + t * Cfg::kQBytes + ch * 16384, &maps.q, q_full, ch * 64, q_base + t * 128, head, b);
    }
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(kv_empty(st), ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full(st), Cfg::kStageBytes);
        const uint32_t kdst = kv_smem + st * Cfg::kStageBytes;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          tma_load_4d(kdst + ch * (BKV * 128), &maps.k, kv_full(st), ch * 64, j * BKV, head, b);
          tma_load_4d(kdst + Cfg::kKBytes + ch * (BKV * 128), &maps.v, kv_full(st), ch * 64, j * BKV, head, b);
        }
      }
      if (++st == Cfg::kStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ===============================
    // warp-uniform loop; every issue block is executed by one elected lane (always the same one, so the commits
    // track the MMAs it issued)
    const uint32_t idesc_s = make_idesc_f16(128, BKV);
    const uint32_t idesc_o = make_idesc_f16(128, dpad, 0, 1);   // B (= V) is MN-major
    const int ks_s = dpad / 16;
    // operand descriptors of tile t / stage st differ from tile 0 / stage 0 by constants (16-byte units)
    const uint64_t q_desc0 = make_desc_kmajor_sw128(base);
    const uint64_t p_desc0 = make_desc_kmajor_sw128(base + Cfg::kPOff);
    const uint64_t k_desc0 = make_desc_kmajor_sw128(kv_smem);
    const uint64_t v_desc0 = make_desc_mnmajor_sw128(kv_smem + Cfg::kKBytes, BKV * 128, 1024);
    const uint32_t idesc_l = make_idesc_f16(128, 16, 0, 1);
    const uint64_t ones_desc = make_desc_mnmajor_sw128(base + Cfg::kOnesOff, BKV * 128, 1024);
    constexpr uint64_t kStage16 = Cfg::kStageBytes >> 4;
    mbar_wait(q_full, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    if (elect_one()) {
      for (int t = 0; t < ntile; ++t) {
        issue_qk_dyn<BKV>(ks_s, tmem_base + t_s_col(t), q_desc0 + t * (Cfg::kQBytes >> 4), k_desc0, idesc_s);
        umma_commit(s_full(t));
      }
    }
    int st = 0;
    uint32_t kv_ph = 0;                        // phase of kv_full(stn) for block j + 1
    for (int j = 0; j < nblk; ++j) {
      const int stn = (st + 1 == Cfg::kStages) ? 0 : st + 1;
      if (stn == 0) kv_ph ^= 1;
      // (1) S_t(j+1) = Q_t K(j+1)^T as soon as softmax_t has pulled S_t(j) into registers (s_free): the next scores
      //     are computed WHILE softmax_t(j) exponentiates, so a softmax warpgroup never waits for its tensor-core work
      if (j + 1 < nblk) {
        mbar_wait(kv_full(stn), kv_ph);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t < ntile) {
            mbar_wait(s_free(t), j & 1);
            tc_fence_after();
            if (elect_one()) {
              issue_qk_dyn<BKV>(ks_s, tmem_base + t_s_col(t), q_desc0 + t * (Cfg::kQBytes >> 4),
                                k_desc0 + stn * kStage16, idesc_s);
              umma_commit(s_full(t));
            }
          }
        }
      }
      // (2) O_t += P_t(j) V(j) once softmax_t(j) has written P_t(j)
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t < ntile) {
          mbar_wait(p_full(t), j & 1);       // softmax_t(j) has written P_t(j)
          tc_fence_after();
          if (elect_one()) {
            ATTN_TRACE(2 + t, j, 0);
            // K loop over the kv rows of this block in steps of 16
            if (kPTmem)
              issue_pv_ts<BKV>(tmem_base + t_o_col(t), tmem_base + t_p_col(t), v_desc0 + st * kStage16, idesc_o, j == 0);
            else
              issue_pv<BKV>(tmem_base + t_o_col(t), p_desc0 + t * (Cfg::kPBytes >> 4), v_desc0 + st * kStage16, idesc_o,
                            j == 0);
            if (kTcSum && kPTmem)                             // L_t += P_t(j) 1: the softmax denominators
              issue_pv_ts<BKV>(tmem_base + t_l_col(t), tmem_base + t_p_col(t), ones_desc, idesc_l, j == 0);
            else if (kTcSum)
              issue_pv<BKV>(tmem_base + t_l_col(t), p_desc0 + t * (Cfg::kPBytes >> 4), ones_desc, idesc_l, j == 0);
            umma_commit(pv_done(t));                          // P_t buffer + O_t free again
            if (t == ntile - 1) umma_commit(kv_empty(st));   // K(j)/V(j) fully consumed by both tiles
            if (j == nblk - 1) umma_commit(o_ready(t));
            ATTN_TRACE(2 + t, j, 1);
          }
        }
      }
      st = stn;
    }
  }
  } else if (SPLIT == 2) {
    // =============================== softmax + epilogue, column-split ===============================
    // warp w: TMEM lane quadrant w & 3, warpgroup g = w >> 2: tile t = g & 1, column half hf = g >> 1
    constexpr int HC = BKV / 2;                    // 64 S columns per thread and block
    const int g = warp >> 2;
    const int t = g & 1, hf = g >> 1;
    if (t < ntile) {
      const int qd = warp & 3;
      const int row = qd * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
      const uint32_t t_s = tmem_base + lane_off + t_s_col(t) + hf * HC;
      const uint32_t t_o = tmem_base + lane_off + t_o_col(t);
      const uint32_t t_p = tmem_base + lane_off + t_p_col(t) + hf * (HC / 2);     // my 32 packed columns of P_t
      const float sl2 = p.scale * 1.4426950408889634f;
      float m_ref = -INFINITY, l_run = 0.f;
      // exchange slots xch[buffer j & 1][tile][half][row] in the (unused: P lives in TMEM) shared-memory P region
      float* const xch = reinterpret_cast<float*>(smem + Cfg::kPOff);
      if ((VAR & kAttnStagger) && t == 1) named_bar_sync(1, 512);    // tile 0's two warpgroups arrive (j = 0)
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(s_full(t), j & 1);
        tc_fence_after();
        float v[HC];
        tmem_ld32(t_s, v);
        tmem_ld32(t_s + 32, v + 32);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free(t));
        const int kv_valid = p.Nk - j * BKV - hf * HC;          // my columns >= kv_valid are padding (last block only)
        if (kv_valid < HC) {
#pragma unroll
          for (int i = 0; i < HC; ++i)
            if (i >= kv_valid) v[i] = -INFINITY;
        }
        float mx4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
        for (int i = 4; i + 8 <= HC; i += 8) {
          mx4[0] = fmax3(mx4[0], v[i], v[i + 1]);
          mx4[1] = fmax3(mx4[1], v[i + 2], v[i + 3]);
          mx4[2] = fmax3(mx4[2], v[i + 4], v[i + 5]);
          mx4[3] = fmax3(mx4[3], v[i + 6], v[i + 7]);
        }
        mx4[0] = fmax3(mx4[0], v[HC - 4], v[HC - 3]);
        mx4[1] = fmax3(mx4[1], v[HC - 2], v[HC - 1]);
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // the row maximum of the whole block: one float per row through shared memory, one named barrier per tile
        float* const slot = xch + (((j & 1) * 2 + t) * 2) * 128;
        slot[hf * 128 + row] = mx;
        named_bar_sync(2 + t, 256);
        mx = fmaxf(mx, slot[(hf ^ 1) * 128 + row]);
        bool pv_waited = (j == 0);
        const bool need = (mx - m_ref) * sl2 > 8.0f;            // same rows, same inputs in both halves: same decision
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_ref, mx);
          const float alpha = fast_exp2((m_ref - m_new) * sl2);
          m_ref = m_new;
          l_run *= alpha;
          if (j > 0) {
            mbar_wait(pv_done(t), (j - 1) & 1);
            tc_fence_after();
            pv_waited = true;
#pragma unroll 1
            for (int c = hf; c < dpad / 16; c += 2) {           // the halves split O's 16-column chunks
              float o[16];
              tmem_ld16(t_o + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] *= alpha;
              tmem_st16(t_o + c * 16, o);
            }
            tmem_st_wait();
          }
        }
        const float mb = m_ref * sl2;
        uint32_t pk[HC / 2];
        const f32x2_t sl2_2 = pack_f32x2(sl2, sl2), nmb_2 = pack_f32x2(-mb, -mb);
        f32x2_t rs2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int u = 0; u < HC / 8; ++u) {
          if ((VAR & kAttnStagger) && j == 0 && t == 0 && ntile == 2 && u == HC / 16) named_bar_arrive(1, 512);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = u * 8 + 2 * e;
            float x0, x1, p0, p1;
            unpack_f32x2(fma_f32x2(pack_f32x2(v[i], v[i + 1]), sl2_2, nmb_2), x0, x1);
            const bool poly = ((kPat >> ((u & 1) * 4 + e)) & 1) != 0;
            if (poly) {
              exp2_poly_x2(x0, x1, p0, p1);
            } else {
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            rs2[e] = add_f32x2(rs2[e], pack_f32x2(p0, p1));
            pk[u * 4 + e] = pack_half2(p0, p1);
          }
        }
        float rs4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float a, b2;
          unpack_f32x2(rs2[e], a, b2);
          rs4[e] = a + b2;
        }
        if (!pv_waited) {
          mbar_wait(pv_done(t), (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st32(t_p, pk);
        tmem_st_wait();
        l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        tc_fence_before();
        mbar_arrive(p_full(t));
      }
      // row sums of the two halves -> one normaliser (same exchange pattern, its own buffer parity)
      float* const slot = xch + (((nblk & 1) * 2 + t) * 2) * 128;
      slot[hf * 128 + row] = l_run;
      named_bar_sync(2 + t, 256);
      l_run += slot[(hf ^ 1) * 128 + row];
      const float inv_l = 1.0f / l_run;
      mbar_wait(o_ready(t), 0);
      tc_fence_after();
      const int q = q_base + t * 128 + row;
      if (hf == 0 && p.lse2 != nullptr && q < p.Nq)
        p.lse2[(static_cast<size_t>(b) * p.heads + head) * p.Nq + q] = m_ref * sl2 + log2f(l_run);
      __half* op = p.out + (static_cast<size_t>(b) * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll 1
      for (int c = hf; c < dpad / 16; c += 2) {
        float o[16];
        tmem_ld16(t_o + c * 16, o);
        tmem_ld_wait();
        if (q < p.Nq) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = c * 16 + u * 8;
            if (col + 8 <= p.d) {
              uint4 w;
              w.x = pack_half2(o[u * 8 + 0] * inv_l, o[u * 8 + 1] * inv_l);
              w.y = pack_half2(o[u * 8 + 2] * inv_l, o[u * 8 + 3] * inv_l);
              w.z = pack_half2(o[u * 8 + 4] * inv_l, o[u * 8 + 5] * inv_l);
              w.w = pack_half2(o[u * 8 + 6] * inv_l, o[u * 8 + 7] * inv_l);
              *reinterpret_cast<uint4*>(op + col) = w;
            }
          }
        }
      }
    }
  } else {
    // =============================== softmax + epilogue (warpgroup t = warp / 4) ===============================
    if (kRegSplit) setmaxnreg_inc<200>();    // 2 x 200 + 104 = 504 = 3 x 168: what the CTA was launched with
    const int t = warp >> 2;
    if (t < ntile) {
      const int qd = warp & 3;                // TMEM lane quadrant
      const int row = qd * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
      const uint32_t t_s = tmem_base + lane_off + t_s_col(t);
      const uint32_t t_o = tmem_base + lane_off + t_o_col(t);
      const float sl2 = p.scale * 1.4426950408889634f;
      float m_ref = -INFINITY, l_run = 0.f;
      const uint32_t p_row = base + Cfg::kPOff + t * Cfg::kPBytes + row * 128;
      const int sw = row & 7;
      if ((VAR & kAttnStagger) && t == 1) named_bar_sync(1, 256);    // ntile == 2 here: tile 0's warpgroup arrives (j = 0)
      const uint32_t t_l = tmem_base + lane_off + t_l_col(t);
      const uint32_t t_p = tmem_base + lane_off + t_p_col(t);
      float v[BKV];                           // S_t(j) of this thread's row
      for (int j = 0; j < nblk; ++j) {
        const bool tr = (lane == 0 && qd == 0);
        if (tr) ATTN_TRACE(t, j, 0);
        const int kv_valid = p.Nk - j * BKV;     // columns >= kv_valid are padding (last block only)
        float mx4[4];                            // 4 independent chains (FMNMX latency, not issue, bounds)
        auto max_first = [&](int hi) {           // FMNMX3: two new elements per instruction; columns [0, hi)
          mx4[0] = v[0]; mx4[1] = v[1]; mx4[2] = v[2]; mx4[3] = v[3];
#pragma unroll
          for (int i = 4; i + 8 <= hi; i += 8) {
            mx4[0] = fmax3(mx4[0], v[i], v[i + 1]);
            mx4[1] = fmax3(mx4[1], v[i + 2], v[i + 3]);
            mx4[2] = fmax3(mx4[2], v[i + 4], v[i + 5]);
            mx4[3] = fmax3(mx4[3], v[i + 6], v[i + 7]);
          }
          mx4[0] = fmax3(mx4[0], v[hi - 4], v[hi - 3]);
          mx4[1] = fmax3(mx4[1], v[hi - 2], v[hi - 1]);
        };
        auto max_more = [&](int lo, int hi) {    // columns [lo, hi), (hi - lo) % 8 == 0
#pragma unroll
          for (int i = lo; i + 8 <= hi; i += 8) {
            mx4[0] = fmax3(mx4[0], v[i], v[i + 1]);
            mx4[1] = fmax3(mx4[1], v[i + 2], v[i + 3]);
            mx4[2] = fmax3(mx4[2], v[i + 4], v[i + 5]);
            mx4[3] = fmax3(mx4[3], v[i + 6], v[i + 7]);
          }
        };
        bool half_done = false;
        if (kEarlyS && j > 0) {
          // the first half of S_t(j) was requested before the previous block's P stores; the second half's TMEM latency
          // hides behind the first half's row maximum
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < BKV / 64; ++c) touch32(v + c * 32);
#pragma unroll
          for (int c = BKV / 64; c < BKV / 32; ++c) tmem_ld32(t_s + c * 32, v + c * 32);
          if (kv_valid >= BKV) {                 // warp-uniform; the padded last block takes the plain path below
            max_first(BKV / 2);
            half_done = true;
          }
          tmem_ld_wait();
#pragma unroll
          for (int c = BKV / 64; c < BKV / 32; ++c) touch32(v + c * 32);
        } else {
          mbar_wait(s_full(t), j & 1);
          tc_fence_after();
          if (tr) ATTN_TRACE(t, j, 1);
#pragma unroll
          for (int c = 0; c < BKV / 32; ++c) tmem_ld32(t_s + c * 32, v + c * 32);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(s_free(t));               // S_t(j) is in registers: the tensor core may overwrite it with S_t(j+1)
        if (tr) ATTN_TRACE(t, j, 2);
        if (kv_valid < BKV) {
#pragma unroll
          for (int i = 0; i < BKV; ++i)
            if (i >= kv_valid) v[i] = -INFINITY;
        }
        if (VAR & kAttnPacked) {
          if (half_done) {
            max_more(BKV / 2, BKV);
          } else {
            max_first(BKV);
          }
        } else {
          mx4[0] = v[0]; mx4[1] = v[1]; mx4[2] = v[2]; mx4[3] = v[3];
#pragma unroll
          for (int i = 4; i < BKV; i += 4) {
            mx4[0] = fmaxf(mx4[0], v[i]);
            mx4[1] = fmaxf(mx4[1], v[i + 1]);
            mx4[2] = fmaxf(mx4[2], v[i + 2]);
            mx4[3] = fmaxf(mx4[3], v[i + 3]);
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // P_t V(j-1) must be complete before O_t is rescaled (rare: lazy rescale) or the P_t buffer is overwritten.
        // The wait is taken as LATE as possible: all exponentials of this block are computed into registers first, so
        // the tensor core's P V work hides behind the MUFU-bound loop instead of stalling it.
        bool pv_waited = (j == 0);
        // lazy rescale: move the reference maximum only when some row of this warp grew by more than 2^8
        const bool need = (mx - m_ref) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = fmaxf(m_ref, mx);
          const float alpha = fast_exp2((m_ref - m_new) * sl2);     // 0 on the first block (m_ref = -inf)
          m_ref = m_new;
          l_run *= alpha;
          if (j > 0) {
            mbar_wait(pv_done(t), (j - 1) & 1);
            tc_fence_after();
            pv_waited = true;
            // P_t V(j-1) is complete (pv_done) and P_t V(j) is not issued before this warpgroup arrives on p_full,
            // so O_t is stable here
#pragma unroll 1
            for (int c = 0; c < dpad / 16; ++c) {
              float o[16];
              tmem_ld16(t_o + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] *= alpha;
              tmem_st16(t_o + c * 16, o);
            }
            if (kTcSum) {
              float o[16];
              tmem_ld16(t_l, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] *= alpha;
              tmem_st16(t_l, o);
            }
            tmem_st_wait();
          }
        }
        if (tr) ATTN_TRACE(t, j, 3);
        const float mb = m_ref * sl2;
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[BKV / 2];

        if ((VAR & kAttnPacked) || kPat != 0) {
          const f32x2_t sl2_2 = pack_f32x2(sl2, sl2), nmb_2 = pack_f32x2(-mb, -mb);
          f32x2_t rs2[4] = {0ull, 0ull, 0ull, 0ull};             // (0.f, 0.f) pairs
#pragma unroll
          for (int u = 0; u < BKV / 8; ++u) {
            if ((VAR & kAttnStagger) && j == 0 && t == 0 && ntile == 2 && u == BKV / 16)
              named_bar_arrive(1, 256);                         // half-way through tile 0's first block: release tile 1
            if (kHalfStore && u == BKV / 16) {
              // first half of the row's probabilities -> P_t (frees 32 registers for the second half).  P_t V(j-1)
              // was issued a whole exponential half-phase ago: the wait is free
              if (!pv_waited) {
                mbar_wait(pv_done(t), (j - 1) & 1);
                tc_fence_after();
                pv_waited = true;
              }
              tmem_st32(t_p, pk);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = u * 8 + 2 * e;
              float x0, x1, p0, p1;
              unpack_f32x2(fma_f32x2(pack_f32x2(v[i], v[i + 1]), sl2_2, nmb_2), x0, x1);
              const bool poly = ((kPat >> ((u & 1) * 4 + e)) & 1) != 0;
              if (poly) {
                exp2_poly_x2(x0, x1, p0, p1);
              } else {
                p0 = fast_exp2(x0);
                p1 = fast_exp2(x1);
              }
              if (!kTcSum) rs2[e] = add_f32x2(rs2[e], pack_f32x2(p0, p1));
              pk[u * 4 + e] = pack_half2(p0, p1);
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a, b2;
            unpack_f32x2(rs2[e], a, b2);
            rs4[e] = a + b2;
          }
        } else {
#pragma unroll
          for (int u = 0; u < BKV / 8; ++u) {
            if ((VAR & kAttnStagger) && j == 0 && t == 0 && ntile == 2 && u == BKV / 16) named_bar_arrive(1, 256);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = u * 8 + 2 * e;
              const float p0 = fast_exp2(v[i] * sl2 - mb);
              const float p1 = fast_exp2(v[i + 1] * sl2 - mb);
              if (!kTcSum) rs4[e] += p0 + p1;
              pk[u * 4 + e] = pack_half2(p0, p1);
            }
          }
        }
        if (!pv_waited) {                     // warp-uniform
          mbar_wait(pv_done(t), (j - 1) & 1);
          tc_fence_after();
        }
        if (kEarlyS && j + 1 < nblk) {        // S_t(j+1) was issued when this block's scores left TMEM: long complete
          mbar_wait(s_full(t), (j + 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < BKV / 64; ++c) tmem_ld32(t_s + c * 32, v + c * 32);     // first half; see the loop head
        }
        if (kPTmem) {                           // this row's 128 probabilities = 64 packed columns of P_t
          if (!kHalfStore) tmem_st32(t_p, pk);  // (else the first 32 columns left half-way through the loop above)
          if (BKV == 128) tmem_st32(t_p + 32, pk + 32);
          tmem_st_wait();
        } else
#pragma unroll
        for (int u = 0; u < BKV / 8; ++u) {
          // 16 B unit (u & 7) of this row lands at ((u & 7) ^ (row & 7)) in the 128B-swizzled 64-column chunk u >> 3
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + (u >> 3) * 16384 + (((u & 7) ^ sw) << 4)),
                       "r"(pk[u * 4 + 0]), "r"(pk[u * 4 + 1]), "r"(pk[u * 4 + 2]), "r"(pk[u * 4 + 3])
                       : "memory");
        }
        l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        if (tr) ATTN_TRACE(t, j, 4);
        if (!kPTmem) fence_proxy_async_shared();   // P (generic-proxy stores) -> visible to the tensor core (async proxy)
        tc_fence_before();
        mbar_arrive(p_full(t));
        if (tr) ATTN_TRACE(t, j, 5);
      }
      mbar_wait(o_ready(t), 0);
      tc_fence_after();
      if (kTcSum) {
        float lt[16];
        tmem_ld16(t_l, lt);
        tmem_ld_wait();
        l_run = lt[0];
      }
      const float inv_l = 1.0f / l_run;
      const int q = q_base + t * 128 + row;
      if (p.lse2 != nullptr && q < p.Nq)      // P was exp2((s - m_ref) * sl2): log2 sum exp2(s * sl2) = m_ref * sl2 + log2(l)
        p.lse2[(static_cast<size_t>(b) * p.heads + head) * p.Nq + q] = m_ref * sl2 + log2f(l_run);
      __half* op = p.out + (static_cast<size_t>(b) * p.Nq + q) * p.ldo + head * p.d;
#pragma unroll 1
      for (int c = 0; c < dpad / 16; ++c) {
        float o[16];
        tmem_ld16(t_o + c * 16, o);
        tmem_ld_wait();
        if (q < p.Nq) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = c * 16 + u * 8;
            if (col + 8 <= p.d) {
              uint4 w;
              w.x = pack_half2(o[u * 8 + 0] * inv_l, o[u * 8 + 1] * inv_l);
              w.y = pack_half2(o[u * 8 + 2] * inv_l, o[u * 8 + 3] * inv_l);
              w.z = pack_half2(o[u * 8 + 4] * inv_l, o[u * 8 + 5] * inv_l);
              w.w = pack_half2(o[u * 8 + 6] * inv_l, o[u * 8 + 7] * inv_l);
              *reinterpret_cast<uint4*>(op + col) = w;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  pdl_launch();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int NCH, int VAR, int SPLIT = 1>
static cudaError_t launch_cfg(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream) {
  using Cfg = AttnCfg<NCH>;
  constexpr int smem_bytes = (VAR & kAttnTcSum) ? Cfg::kSmemBytesOnes : Cfg::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tcgen05_kernel<NCH, VAR, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((p.Nq + 255) / 256, p.heads, p.B);
  return launch_pdl(attention_tcgen05_kernel<NCH, VAR, SPLIT>, grid, dim3(attn_threads<SPLIT>()), smem_bytes, stream, maps, p);
}

int attention_bkv(int d) { return d <= 64 ? 128 : 64; }

// softmax variant (bits above): the default is the measured optimum; UNIB200_ATTN_VARIANT overrides it for A/B runs
static int attention_variant(int nch) {
  static const int v = getenv("UNIB200_ATTN_VARIANT") ? atoi(getenv("UNIB200_ATTN_VARIANT")) : -1;
  if (v >= 0 && (nch == 1 || v < 16)) return v;      // variants >= 16 exist for head dims <= 64 only
  static const int v2 = getenv("UNIB200_ATTN_VARIANT2") ? atoi(getenv("UNIB200_ATTN_VARIANT2")) : 65;
  return nch == 1 ? UNIB_ATTN_DEFAULT_VARIANT : nch == 2 ? v2 : 0;    // head dims 65..128: packed softmax, P in TMEM (27.4 -> 24.9 us at 1024 x 1024, d = 80)
}

template <int NCH>
static cudaError_t launch_var(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream) {
  switch (attention_variant(NCH)) {
    case 0: return launch_cfg<NCH, 0>(maps, p, stream);
    case 1: return launch_cfg<NCH, 1>(maps, p, stream);
    case 2: return launch_cfg<NCH, 2>(maps, p, stream);
    case 3: return launch_cfg<NCH, 3>(maps, p, stream);
    case 5: return launch_cfg<NCH, 5>(maps, p, stream);
    case 7: return launch_cfg<NCH, 7>(maps, p, stream);
    case 9: return launch_cfg<NCH, 9>(maps, p, stream);
    case 11: return launch_cfg<NCH, 11>(maps, p, stream);
  }
  if constexpr (NCH == 2) {
    if (attention_variant(NCH) == 64) return launch_cfg<2, 64>(maps, p, stream);
    if (attention_variant(NCH) == 65) return launch_cfg<2, 65>(maps, p, stream);
  }
  if constexpr (NCH == 1) {                  // round-2 experiments (head dims <= 64)
    // UNIB200_ATTN_SPLIT=1: column-split softmax (16 softmax warps) for long key sequences
    static const int split = getenv("UNIB200_ATTN_SPLIT") ? atoi(getenv("UNIB200_ATTN_SPLIT")) : 0;
    if (split && p.Nk >= 512) {
      switch (attention_variant(NCH)) {
        case 67: return launch_cfg<1, 67, 2>(maps, p, stream);
        case 75: return launch_cfg<1, 75, 2>(maps, p, stream);
        case 43075: return launch_cfg<1, 67 | (0xA8 << 8), 2>(maps, p, stream);
        default: return launch_cfg<1, 71, 2>(maps, p, stream);
      }
    }
    switch (attention_variant(NCH)) {
      case 23: return launch_cfg<1, 23>(maps, p, stream);                    // 7 + tensor-core row sums
      case 27: return launch_cfg<1, 27>(maps, p, stream);                    // tc sums, polynomial on every 2nd pair
      case 43011: return launch_cfg<1, 3 | (0xA8 << 8)>(maps, p, stream);    // polynomial on 3 of 8 pairs
      case 43027: return launch_cfg<1, 19 | (0xA8 << 8)>(maps, p, stream);   // + tc sums
      case 67: return launch_cfg<1, 67>(maps, p, stream);                    // P in tensor memory, no polynomial
      case 71: return launch_cfg<1, 71>(maps, p, stream);                    // 7 + P in tensor memory
      case 199: return launch_cfg<1, 199>(maps, p, stream);                  // + half-way P store
      case 203: return launch_cfg<1, 203>(maps, p, stream);                  // polynomial on every 2nd pair
      case 43203: return launch_cfg<1, 195 | (0xA8 << 8)>(maps, p, stream);  // polynomial on 3 of 8 pairs
      // P in tensor memory + tensor-core row sums: TMEM only has room for head dims <= 48
      case 87: return p.d <= 48 ? launch_cfg<1, 87>(maps, p, stream) : launch_cfg<1, 71>(maps, p, stream);
      case 215: return p.d <= 48 ? launch_cfg<1, 215>(maps, p, stream) : launch_cfg<1, 199>(maps, p, stream);
      case 219: return p.d <= 48 ? launch_cfg<1, 219>(maps, p, stream) : launch_cfg<1, 203>(maps, p, stream);
      case 43091: return p.d <= 48 ? launch_cfg<1, 83 | (0xA8 << 8)>(maps, p, stream)
                                   : launch_cfg<1, 195 | (0xA8 << 8)>(maps, p, stream);
      case 43219: return p.d <= 48 ? launch_cfg<1, 211 | (0xA8 << 8)>(maps, p, stream)
                                   : launch_cfg<1, 195 | (0xA8 << 8)>(maps, p, stream);
    }
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_attention(const AttnMaps& maps, const AttnParams& p, cudaStream_t stream) {
  if (p.d % 8 != 0 || p.d > 192 || p.d < 8) return cudaErrorInvalidValue;
  if (p.d <= 64) return launch_var<1>(maps, p, stream);
  if (p.d <= 128) return launch_var<2>(maps, p, stream);
  return launch_var<3>(maps, p, stream);
}

}  // namespace unib
