// C ABI (include/unib200.h): argument validation, TMA tensor-map encoding, split-K planning, program recording and
// CUDA-graph replay.  No torch types, no exceptions across the boundary.
#include <cuda.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include <functional>
#include <string>
#include <vector>

#include "../../include/unib200.h"
#include "attention_bwd_sm100.cuh"
#include "attention_sm100.cuh"
#include "elementwise.cuh"
#include "gemm_sm100.cuh"
#include "wgrad_sm100.cuh"

using namespace unib;

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}
int fail_cuda(const char* what, cudaError_t e) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return -2;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// fp16 tensor map, SWIZZLE_128B, zero OOB fill.  dims/strides innermost first; strides[i] is the byte stride of dim i+1.
bool encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, std::string* why, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { *why = "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)"; return false; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) { *why = "tensor base not 16-byte aligned"; return false; }
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16 != 0) { *why = "tensor stride not a multiple of 16 bytes"; return false; }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *why = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

long long* g_attn_trace = nullptr;
int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

typedef std::function<cudaError_t(cudaStream_t)> Op;

// what an op is, for per-kind accounting of a recorded program (unib200_program_op_info / _profile)
struct OpInfo {
  int kind = UNIB200_OP_OTHER;
  double flops = 0.0;     // algorithmic (2 * MAC, unpadded shapes)
  double bytes = 0.0;     // algorithmic HBM bytes (each operand read once, result written once)
  int launches = 1;
  std::string desc;       // human-readable shape summary (tools/profile_ops.py)
};

}  // namespace

constexpr int kMaxLanes = 4;
constexpr int kBarrierLane = -1;

struct unib200_program {
  std::vector<Op> ops;
  std::vector<OpInfo> info;
  std::vector<int> lane;          // execution lane of each op; kBarrierLane marks a fork/join point
  int cur_lane = 0;
  int launches = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  // lanes 1.. run on side streams owned by the program; events are handed out in order and reused run to run
  cudaStream_t lane_stream[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> events;
};

namespace {
// measurement aid: UNIB200_SKIP_KINDS=<bitmask of UNIB200_OP_*> records no-ops for those op kinds, so the marginal
// cost of an op class inside the real multi-lane step graph can be read off a bench run (results are garbage)
int skip_mask() {
  static int m = -1;
  if (m < 0) {
    const char* e = getenv("UNIB200_SKIP_KINDS");
    m = e ? atoi(e) : 0;
  }
  return m;
}

int submit(unib200_program* prog, Op op, int launches, void* stream, const char* what, int kind = UNIB200_OP_OTHER,
           double flops = 0.0, double bytes = 0.0, const std::string& desc = std::string()) {
  if (prog && (skip_mask() & (1 << kind))) {
    op = [](cudaStream_t) { return cudaSuccess; };
    launches = 0;
  }
  if (prog) {
    prog->ops.push_back(std::move(op));
    prog->lane.push_back(prog->cur_lane);
    OpInfo oi;
    oi.kind = kind; oi.flops = flops; oi.bytes = bytes; oi.launches = launches;
    oi.desc = desc.empty() ? std::string(what) : desc;
    prog->info.push_back(oi);
    prog->launches += launches;
    return 0;
  }
  cudaError_t e = op(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(what, e);
  return 0;
}
}  // namespace

extern "C" {

int unib200_version(void) { return UNIB200_VERSION; }
const char* unib200_last_error(void) { return g_err.c_str(); }

int unib200_device_info(int* sms, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail_cuda("cudaGetDevice", e);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return fail_cuda("cudaGetDeviceProperties", e);
  if (sms) *sms = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

unib200_program* unib200_program_create(void) { return new (std::nothrow) unib200_program(); }

void unib200_program_destroy(unib200_program* prog) {
  if (!prog) return;
  if (prog->exec) cudaGraphExecDestroy(prog->exec);
  if (prog->graph) cudaGraphDestroy(prog->graph);
  for (int l = 1; l < kMaxLanes; ++l)
    if (prog->lane_stream[l]) cudaStreamDestroy(prog->lane_stream[l]);
  for (cudaEvent_t e : prog->events) cudaEventDestroy(e);
  delete prog;
}

int unib200_program_set_lane(unib200_program* prog, int lane) {
  if (!prog) return fail("null program");
  if (lane < 0 || lane >= kMaxLanes) return fail("program_set_lane: lane must be in [0, 4)");
  prog->cur_lane = lane;
  return 0;
}

int unib200_program_barrier(unib200_program* prog) {
  if (!prog) return fail("null program");
  prog->ops.push_back(Op());
  prog->lane.push_back(kBarrierLane);
  OpInfo oi;
  oi.launches = 0;
  oi.desc = "barrier";
  prog->info.push_back(oi);
  return 0;
}

int unib200_program_num_launches(const unib200_program* prog) { return prog ? prog->launches : 0; }

// Lane 0 ops are launched on `stream`; ops of lanes 1.. on the program's side streams, forked from `stream` at the last
// barrier (or the start of the program) and joined back at the next barrier (or the end).  Works identically in
// eager mode and under stream capture, where the forks/joins become parallel branches of the CUDA graph.
int unib200_program_run(unib200_program* prog, void* stream) {
  if (!prog) return fail("null program");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  size_t next_ev = 0;
  cudaError_t e = cudaSuccess;
  auto event = [&]() -> cudaEvent_t {
    if (next_ev == prog->events.size()) {
      cudaEvent_t ev = nullptr;
      e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) return nullptr;
      prog->events.push_back(ev);
    }
    return prog->events[next_ev++];
  };
  bool active[kMaxLanes] = {false, false, false, false};
  cudaEvent_t fork = nullptr;        // recorded on `s` lazily, when the first side-lane op after a barrier shows up
  auto join = [&]() -> cudaError_t {
    for (int l = 1; l < kMaxLanes; ++l) {
      if (!active[l]) continue;
      cudaEvent_t ev = event();
      if (!ev) return e;
      if ((e = cudaEventRecord(ev, prog->lane_stream[l])) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(s, ev, 0)) != cudaSuccess) return e;
      active[l] = false;
    }
    fork = nullptr;
    return cudaSuccess;
  };
  // fork events must be recorded at the barrier itself (not later, after more lane-0 work was queued)
  bool has_side = false;
  for (int l : prog->lane) has_side = has_side || l > 0;
  auto record_fork = [&]() -> cudaError_t {
    if (!has_side) return cudaSuccess;
    fork = event();
    if (!fork) return e;
    return cudaEventRecord(fork, s);
  };
  if ((e = record_fork()) != cudaSuccess) return fail_cuda("program fork", e);
  for (size_t i = 0; i < prog->ops.size(); ++i) {
    const int l = prog->lane[i];
    if (l == kBarrierLane) {
      if ((e = join()) != cudaSuccess) return fail_cuda("program barrier (join)", e);
      if ((e = record_fork()) != cudaSuccess) return fail_cuda("program barrier (fork)", e);
      continue;
    }
    cudaStream_t target = s;
    if (l > 0) {
      if (!prog->lane_stream[l]) {
        e = cudaStreamCreateWithFlags(&prog->lane_stream[l], cudaStreamNonBlocking);
        if (e != cudaSuccess) return fail_cuda("cudaStreamCreate (lane)", e);
      }
      if (!active[l]) {
        if ((e = cudaStreamWaitEvent(prog->lane_stream[l], fork, 0)) != cudaSuccess) return fail_cuda("lane fork", e);
        active[l] = true;
      }
      target = prog->lane_stream[l];
    }
    e = prog->ops[i](target);
    if (e != cudaSuccess) return fail_cuda(("program op " + std::to_string(i)).c_str(), e);
  }
  if ((e = join()) != cudaSuccess) return fail_cuda("program join", e);
  return 0;
}

int unib200_program_num_ops(const unib200_program* prog) { return prog ? static_cast<int>(prog->ops.size()) : 0; }

int unib200_program_op_info(const unib200_program* prog, int i, int* kind, double* flops, double* bytes,
                            int* launches) {
  if (!prog || i < 0 || i >= static_cast<int>(prog->info.size())) return fail("op_info: bad program / index");
  const OpInfo& oi = prog->info[i];
  if (kind) *kind = oi.kind;
  if (flops) *flops = oi.flops;
  if (bytes) *bytes = oi.bytes;
  if (launches) *launches = oi.launches;
  return 0;
}

const char* unib200_program_op_desc(const unib200_program* prog, int i) {
  if (!prog || i < 0 || i >= static_cast<int>(prog->info.size())) return "";
  return prog->info[i].desc.c_str();
}

// Replays the program `iters` times with a CUDA event pair around every op (on `stream`, the launching stream) and
// returns the mean device time of each op in milliseconds.  Host-synchronous; for measurement only.
int unib200_program_profile(unib200_program* prog, void* stream, int iters, float* ms_out) {
  if (!prog || !ms_out || iters < 1) return fail("program_profile: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n = prog->ops.size();
  std::vector<cudaEvent_t> ev(2 * n);
  for (size_t i = 0; i < 2 * n; ++i) {
    cudaError_t e = cudaEventCreate(&ev[i]);
    if (e != cudaSuccess) return fail_cuda("cudaEventCreate", e);
  }
  std::vector<double> acc(n, 0.0);
  int rc = 0;
  for (int it = 0; it < iters && rc == 0; ++it) {
    for (size_t i = 0; i < n; ++i) {
      cudaEventRecord(ev[2 * i], s);
      cudaError_t e = prog->lane[i] == kBarrierLane ? cudaSuccess : prog->ops[i](s);   // lanes are serialised here
      cudaEventRecord(ev[2 * i + 1], s);
      if (e != cudaSuccess) { rc = fail_cuda(("program op " + std::to_string(i)).c_str(), e); break; }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == 0 && e != cudaSuccess) rc = fail_cuda("cudaStreamSynchronize", e);
    if (rc != 0) break;
    for (size_t i = 0; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
      acc[i] += ms;
    }
  }
  for (size_t i = 0; i < 2 * n; ++i) cudaEventDestroy(ev[i]);
  if (rc != 0) return rc;
  for (size_t i = 0; i < n; ++i) ms_out[i] = static_cast<float>(acc[i] / iters);
  return 0;
}

int unib200_program_graph_instantiate(unib200_program* prog, void* stream) {
  if (!prog) return fail("null program");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (prog->exec) { cudaGraphExecDestroy(prog->exec); prog->exec = nullptr; }
  if (prog->graph) { cudaGraphDestroy(prog->graph); prog->graph = nullptr; }
  cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) return fail_cuda("cudaStreamBeginCapture", e);
  int rc = unib200_program_run(prog, stream);
  cudaGraph_t g = nullptr;
  e = cudaStreamEndCapture(s, &g);
  if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess) return fail_cuda("cudaStreamEndCapture", e);
  prog->graph = g;
  e = cudaGraphInstantiate(&prog->exec, g, 0);
  if (e != cudaSuccess) return fail_cuda("cudaGraphInstantiate", e);
  return 0;
}

int unib200_program_graph_launch(unib200_program* prog, void* stream) {
  if (!prog || !prog->exec) return fail("program has no instantiated graph");
  cudaError_t e = cudaGraphLaunch(prog->exec, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda("cudaGraphLaunch", e);
  return 0;
}

void unib200_set_pdl(int mode) { unib::g_pdl_enabled = mode < 0 ? 0 : mode; }

void unib200_debug_set_trace(void* dev_buf) { g_attn_trace = static_cast<long long*>(dev_buf); }

int unib200_pick_bn(int N, int flags) { return gemm_pick_bn(N, flags); }

size_t unib200_packed_k(int nseg, const unib200_seg* seg) {
  size_t k = 0;
  for (int i = 0; i < nseg; ++i) {
    const int taps = seg[i].kind == UNIB200_SEG_1x1 ? 1 : seg[i].kind == UNIB200_SEG_UP2x2 ? 4 : 9;
    k += static_cast<size_t>(taps) * ((seg[i].C + 63) / 64) * 64;
  }
  return k;
}

}  // extern "C"

namespace {
struct PreparedGemm {
  GemmMaps maps;
  GemmParams p;
  int bn = 0, launches = 1;
  double flops = 0.0, bytes = 0.0;
  std::string desc;
};

// validation, tensor maps, tiling / split-K / epilogue-mode planning of one conv_gemm descriptor
int prepare_gemm(const unib200_gemm_desc* d, PreparedGemm* out) {
  if (!d) return fail("null desc");
  if (d->M <= 0 || d->N <= 0 || d->nseg < 1 || d->nseg > kMaxSeg) return fail("conv_gemm: bad M/N/nseg");
  const bool linear = (d->H == 0 || d->W == 0);
  GemmMaps& maps = out->maps;
  memset(&maps, 0, sizeof(maps));
  GemmParams& p = out->p;
  memset(&p, 0, sizeof(p));
  p.M = d->M;
  p.N = d->N;
  std::string why;
  int bw = 128, bh = 1, bb = 1;
  if (linear) {
    p.W = 1 << 30;
    p.H = 1;
    p.rows_per_batch = d->B > 0 ? d->M / d->B : d->M;
    if (p.rows_per_batch <= 0) p.rows_per_batch = d->M;
  } else {
    if (d->B * d->H * d->W != d->M) return fail("conv_gemm: M != B*H*W");
    auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
    // W > 128 (VAE resolutions): a 128-pixel tile is then a piece of one image row, box {64, 128, 1, 1}
    if (!pow2(d->W) || !pow2(d->H)) return fail("conv_gemm: H and W must be powers of two");
    p.W = d->W;
    p.H = d->H;
    p.rows_per_batch = d->H * d->W;
    bw = d->W < 128 ? d->W : 128;
    bh = 128 / bw;
    if (bh > d->H) bh = d->H;
    bb = 128 / (bw * bh);
  }
  int nmaps = 0, total_kb = 0;
  p.nseg = d->nseg;
  for (int i = 0; i < d->nseg; ++i) {
    const unib200_seg& s = d->seg[i];
    if (s.C <= 0 || s.ld < s.C || s.ld % 8 != 0) return fail("conv_gemm: bad segment C/ld (ld must be a multiple of 8)");
    if (s.kind < UNIB200_SEG_1x1 || s.kind > UNIB200_SEG_UP2x2) return fail("conv_gemm: bad segment kind");
    if (s.kind == UNIB200_SEG_UP2x2 && d->nseg != 1) return fail("conv_gemm: SEG_UP2x2 must be the only segment");
    if (linear && s.kind != UNIB200_SEG_1x1) return fail("conv_gemm: linear A only supports 1x1 segments");
    p.seg[i].tmap = nmaps;
    p.seg[i].kind = s.kind;
    p.seg[i].nkb = (s.C + 63) / 64;
    p.seg[i].ntaps = s.kind == UNIB200_SEG_1x1 ? 1 : s.kind == UNIB200_SEG_UP2x2 ? 4 : 9;
    total_kb += p.seg[i].ntaps * p.seg[i].nkb;
    const uint64_t pix = static_cast<uint64_t>(s.ld) * 2;
    if (linear) {
      if (nmaps + 1 > kMaxAMaps) return fail("conv_gemm: too many tensor maps");
      const uint64_t dims[4] = {static_cast<uint64_t>(s.C), static_cast<uint64_t>(d->M), 1, 1};
      const uint64_t st[3] = {pix, pix * d->M, pix * d->M};
      const uint32_t box[4] = {64, 128, 1, 1};
      if (!encode_map(&maps.a[nmaps++], s.ptr, 4, dims, st, box, &why)) return fail("conv_gemm A map: " + why);
    } else if (s.kind == UNIB200_SEG_3x3_S2 || s.kind == UNIB200_SEG_3x3_S2P0) {
      if (nmaps + 4 > kMaxAMaps) return fail("conv_gemm: too many tensor maps");
      const int Hi = 2 * d->H, Wi = 2 * d->W;   // source image is twice the output size
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) {
          const char* base = static_cast<const char*>(s.ptr) + (static_cast<uint64_t>(ph) * Wi + pw) * pix;
          const uint64_t dims[4] = {static_cast<uint64_t>(s.C), static_cast<uint64_t>(d->W),
                                    static_cast<uint64_t>(d->H), static_cast<uint64_t>(d->B)};
          const uint64_t st[3] = {2 * pix, 2 * pix * Wi, pix * Wi * Hi};
          const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bb)};
          if (!encode_map(&maps.a[nmaps++], base, 4, dims, st, box, &why)) return fail("conv_gemm A map: " + why);
        }
    } else {
      if (nmaps + 1 > kMaxAMaps) return fail("conv_gemm: too many tensor maps");
      const uint64_t dims[4] = {static_cast<uint64_t>(s.C), static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                                static_cast<uint64_t>(d->B)};
      const uint64_t st[3] = {pix, pix * d->W, pix * d->W * d->H};
      const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bb)};
      if (!encode_map(&maps.a[nmaps++], s.ptr, 4, dims, st, box, &why)) return fail("conv_gemm A map: " + why);
    }
  }
  for (int i = nmaps; i < kMaxAMaps; ++i) maps.a[i] = maps.a[0];
  p.total_kb = total_kb;
  const int bn = gemm_pick_bn(d->N, d->flags);
  if ((d->flags & UNIB200_EPI_GEGLU) && (bn == 0 || d->N % bn != 0))
    return fail("conv_gemm: GEGLU needs N divisible by 64 (whole value/gate sub-tiles)");
  {
    const uint64_t ktot = static_cast<uint64_t>(total_kb) * 64;
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->N)};
    const uint64_t st[1] = {ktot * 2};
    // CTA pairs (cta_group::2: each CTA loads half of the B tile) for the mainloop-bound layers: at least two 128-row
    // tiles and a long K loop.  Measured (tools/bench_gemm.py): conv3x3 +14..18 %, but short-K 1x1 layers lose ~10 %
    // to the pair's cluster synchronisation, so those stay on single CTAs.
    static const int pair_min_kb = getenv("UNIB200_PAIR_MIN_KB") ? atoi(getenv("UNIB200_PAIR_MIN_KB")) : 24;
    static const int geglu_pair = getenv("UNIB200_GEGLU_PAIR") ? atoi(getenv("UNIB200_GEGLU_PAIR")) : 0;
    if (d->flags & UNIB200_EPI_GEGLU)
      p.cg = (geglu_pair && bn == 256 && d->M >= 2 * kBM) ? 2 : 1;
    else
      p.cg = (pair_min_kb > 0 && gemm_pair_supported(bn) && d->M > kBM && total_kb >= pair_min_kb) ? 2 : 1;
    const uint32_t box[2] = {64, static_cast<uint32_t>(bn / p.cg)};
    if (!encode_map(&maps.b, d->weight, 2, dims, st, box, &why)) return fail("conv_gemm B map: " + why);
  }
  p.m_tiles = (d->M + kBM * p.cg - 1) / (kBM * p.cg);
  p.n_tiles = (d->N + bn - 1) / bn;
  {
    auto recip = [](unsigned long long dv) { return (1ull << 40) / dv + 1ull; };
    auto ilog2 = [](int x) { int s = 0; while ((1 << s) < x) ++s; return s; };
    if (static_cast<long long>(p.m_tiles) * p.n_tiles * 16 >= (1 << 20)) return fail("conv_gemm: too many tiles");
    p.mul_tiles = recip(static_cast<unsigned long long>(p.m_tiles) * p.n_tiles);
    p.mul_ntiles = recip(static_cast<unsigned long long>(p.n_tiles));
    p.w_shift = ilog2(p.W);
    p.h_shift = ilog2(p.H);
    const int rpb = p.rows_per_batch;
    p.rpb_shift = (rpb > 0 && (rpb & (rpb - 1)) == 0) ? ilog2(rpb) : -1;
  }
  p.up_shift = -1;
  const bool up = d->seg[0].kind == UNIB200_SEG_UP2x2;
  if (up) {
    const int cout = d->N / 4;
    const int tpp = cout / bn;                       // N tiles per output parity
    if (d->N % 4 || cout % bn || tpp < 1 || (tpp & (tpp - 1)))
      return fail("conv_gemm: SEG_UP2x2 needs N = 4 * Cout with Cout a power-of-two multiple of the N tile");
    if (d->res || d->rowstats_out || d->ln_rowstats || d->flags != 0)
      return fail("conv_gemm: SEG_UP2x2 supports bias (and GroupNorm statistics) only");
    if (d->gn_part && d->gn_rows != 128) return fail("conv_gemm: SEG_UP2x2 GroupNorm statistics need gn_rows = 128");
    p.up_cout = cout;
    p.up_shift = 0;
    while ((1 << p.up_shift) < tpp) ++p.up_shift;
  }
  p.bias = d->bias;
  p.bias_bstride = d->bias_bstride;
  p.bias_step = d->bias_step;
  p.bias_step_stride = d->bias_step_stride;
  p.res = static_cast<const __half*>(d->res);
  p.ldr = d->ldr;
  p.out = d->out;
  p.ldc = d->ldc;
  p.flags = d->flags;
  p.axpby = d->axpby;
  p.axpby_step = d->axpby_step;
  p.aux = d->aux;
  p.aux_out = d->aux_out;
  p.axpby_n0 = d->axpby_first_channel;
  p.trace = g_attn_trace;
  p.pdl_early = unib::g_pdl_enabled == 2 ? 1 : 0;
  p.rowstats_out = d->rowstats_out;
  p.ln_rowstats = d->ln_rowstats;
  p.ln_wsum = d->ln_wsum;
  p.ln_parts = d->ln_parts;
  p.ln_eps = d->ln_eps;
  p.ln_inv_c = d->ln_C > 0 ? 1.0f / static_cast<float>(d->ln_C) : 0.f;
  if (d->ln_rowstats && (!d->ln_wsum || d->ln_parts < 1 || d->ln_C < 1))
    return fail("conv_gemm: ln_rowstats needs ln_wsum, ln_parts >= 1 and ln_C >= 1");
  if ((d->rowstats_out || d->ln_rowstats) && (d->flags & UNIB200_EPI_OUT_NCHW))
    return fail("conv_gemm: row statistics / LayerNorm fold need the NHWC fp16 epilogue");
  if (!(d->flags & UNIB200_EPI_OUT_NCHW)) {
    if (d->ldc % 8 != 0) return fail("conv_gemm: ldc must be a multiple of 8");
    if (d->res && d->ldr % 8 != 0) return fail("conv_gemm: ldr must be a multiple of 8");
  }
  if ((d->flags & UNIB200_EPI_AXPBY) && (!(d->flags & UNIB200_EPI_OUT_NCHW) || !d->axpby || !d->aux || !d->aux_out))
    return fail("conv_gemm: EPI_AXPBY needs EPI_OUT_NCHW, axpby, aux and aux_out");
  // split-K: fill the machine when the output has too few tiles (tiny-M layers are weight-bandwidth bound)
  int splits = d->splits;
  const int tiles = p.m_tiles * p.n_tiles * p.cg;       // CTAs one pass over the output keeps busy
  const int sms = num_sms();
  const bool can_split = d->partial != nullptr && !(d->flags & (UNIB200_EPI_GEGLU)) && !d->rowstats_out &&
                         !d->ln_rowstats && !d->gn_part && !up;
  if (splits <= 0) {
    splits = 1;
    if (can_split && tiles * 2 <= sms) {
      splits = sms / tiles;
      // a split costs a second launch (finalize) + an fp32 round trip of the tile: only worth it when each split
      // still has a long K loop (>= 24 K blocks ~ 7 us of mainloop; measured optimum of the step); short-K layers run unsplit and leave the idle
      // SMs to the other lane of the step graph
      static const int min_kb = getenv("UNIB200_SPLIT_MIN_KB") ? atoi(getenv("UNIB200_SPLIT_MIN_KB")) : 24;
      const int max_by_k = total_kb / min_kb > 0 ? total_kb / min_kb : 1;
      if (splits > max_by_k) splits = max_by_k;
      if (splits > 16) splits = 16;
    }
  }
  if (splits > 1) {
    if (!can_split) return fail("conv_gemm: split-K requested without workspace / with GEGLU");
    const size_t need = static_cast<size_t>(splits) * d->M * d->N * sizeof(float);
    if (need > d->partial_bytes) {
      splits = static_cast<int>(d->partial_bytes / (static_cast<size_t>(d->M) * d->N * sizeof(float)));
      if (splits < 1) splits = 1;
    }
    if (splits > total_kb) splits = total_kb;
  }
  p.splits = splits;
  p.partial = d->partial;
  // epilogue mode (csrc/gemm_sm100.cu GemmMode): the register -> 256-bit-store epilogue needs an NHWC fp16 output whose
  // rows are 32 B aligned and whose width is a multiple of 32; everything else takes the direct-store epilogue
  {
    auto al32 = [](const void* ptr, int ld) { return (reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 16 == 0; };
    const int n_out = (d->flags & UNIB200_EPI_GEGLU) ? d->N / 2 : d->N;
    const bool vec_ok = !(d->flags & (UNIB200_EPI_OUT_NCHW | UNIB200_EPI_SILU)) && splits == 1 && d->out != nullptr &&
                        n_out % 32 == 0 && d->N % 32 == 0 && al32(d->out, d->ldc) && (!d->res || al32(d->res, d->ldr));
    if (d->flags & UNIB200_EPI_GEGLU) {
      if (!vec_ok || d->res) return fail("conv_gemm: GEGLU needs an aligned NHWC fp16 output (N / 2 a multiple of 32), no "
                                         "residual, no split-K / NCHW");
      p.mode = 1;
    } else {
      p.mode = vec_ok ? 0 : 2;
    }
    if (p.mode == 2 && (d->rowstats_out || d->ln_rowstats || d->gn_part))
      return fail("conv_gemm: row statistics / LayerNorm fold / GroupNorm statistics need the vector epilogue (aligned "
                  "NHWC fp16 output, N a multiple of 32, no split-K)");
    if (up && p.mode != 0) return fail("conv_gemm: SEG_UP2x2 needs the vector epilogue (aligned NHWC fp16 output)");
    if (d->gn_part) {
      if (p.mode != 0) return fail("conv_gemm: GroupNorm statistics cannot be combined with GEGLU");
      if (d->gn_gran < 2 || d->gn_gran % 2 || bn % d->gn_gran || d->N % d->gn_gran || (up && (d->N / 4) % d->gn_gran))
        return fail("conv_gemm: gn_gran must be even and divide both N and the N tile");
      if ((d->gn_rows != 32 && d->gn_rows != 64 && d->gn_rows != 128) || d->M % d->gn_rows)
        return fail("conv_gemm: gn_rows must be 32, 64 or 128 and divide M");
      p.gn_part = d->gn_part;
      p.gn_gran = d->gn_gran;
      p.gn_rows = d->gn_rows;
    }
    if (p.mode == 2 && !(d->flags & UNIB200_EPI_OUT_NCHW) &&
        ((reinterpret_cast<uintptr_t>(d->out) & 15) || (d->res && (reinterpret_cast<uintptr_t>(d->res) & 15))))
      return fail("conv_gemm: out / res must be 16-byte aligned");
  }
  double kreal = 0.0, a_bytes = 0.0;
  for (int i = 0; i < d->nseg; ++i) {
    const int taps = d->seg[i].kind == UNIB200_SEG_1x1 ? 1 : d->seg[i].kind == UNIB200_SEG_UP2x2 ? 4 : 9;
    kreal += static_cast<double>(taps) * d->seg[i].C;
    // each source pixel is read once algorithmically (a stride-2 conv reads 4x the output pixels)
    a_bytes += 2.0 * d->seg[i].C * d->M *
               ((d->seg[i].kind == UNIB200_SEG_3x3_S2 || d->seg[i].kind == UNIB200_SEG_3x3_S2P0) ? 4.0 : 1.0);
  }
  const double n_out = (d->flags & UNIB200_EPI_GEGLU) ? d->N / 2.0 : d->N;
  const double flops = 2.0 * d->M * d->N * kreal;
  const double bytes = a_bytes + 2.0 * d->N * kreal + ((d->flags & UNIB200_EPI_OUT_F32) ? 4.0 : 2.0) * d->M * n_out +
                       (d->res ? 2.0 * d->M * d->N : 0.0);
  std::string desc = "gemm M=" + std::to_string(d->M) + " N=" + std::to_string(d->N) + " K=" +
                     std::to_string(static_cast<long long>(kreal)) + " bn=" + std::to_string(bn) + " splits=" +
                     std::to_string(splits) + " tiles=" + std::to_string(tiles) + " segs=";
  for (int i = 0; i < d->nseg; ++i)
    desc += std::string(i ? "+" : "") + (d->seg[i].kind == UNIB200_SEG_1x1 ? "1x1:" : d->seg[i].kind == UNIB200_SEG_3x3 ? "3x3:" : d->seg[i].kind == UNIB200_SEG_3x3_S2 ? "3x3s2:" : d->seg[i].kind == UNIB200_SEG_UP2x2 ? "up2x2:" : "3x3s2p0:") +
            std::to_string(d->seg[i].C);
  if (d->flags) desc += " flags=" + std::to_string(d->flags);
  out->bn = bn;
  out->launches = splits > 1 ? 2 : 1;
  out->flops = flops;
  out->bytes = bytes;
  out->desc = desc;
  return 0;
}
}  // namespace

extern "C" {

int unib200_conv_gemm(unib200_program* prog, const unib200_gemm_desc* d, void* stream) {
  PreparedGemm g;
  int rc = prepare_gemm(d, &g);
  if (rc != 0) return rc;
  const int sms = num_sms();
  const int bn = g.bn;
  const GemmMaps maps = g.maps;
  const GemmParams p = g.p;
  Op op = [maps, p, bn, sms](cudaStream_t s) { return launch_gemm(maps, p, bn, sms, s); };
  // measurement aid (see skip_mask): UNIB200_SKIP_GEMM=<bitmask> drops one GEMM class from recorded programs --
  // 1 short-K single-CTA tiles, 2 GEGLU, 4 split-K (+ finalize), 8 long-K CTA pairs, 16 long-K single CTA
  static const int skip_cls = getenv("UNIB200_SKIP_GEMM") ? atoi(getenv("UNIB200_SKIP_GEMM")) : 0;
  if (prog && skip_cls) {
    const int cls = (d->flags & UNIB200_EPI_GEGLU) ? 2 : p.splits > 1 ? 4 : p.cg == 2 ? 8 : p.total_kb < 24 ? 1 : 16;
    if (skip_cls & cls) {
      op = [](cudaStream_t) { return cudaSuccess; };
      g.launches = 0;
    }
  }
  return submit(prog, std::move(op), g.launches, stream, "conv_gemm", UNIB200_OP_GEMM, g.flops, g.bytes, g.desc);
}

// G = 2 grouped launch: two GEMMs of identical shape (M, N, one 1x1 / linear segment of the same C, same flags) in ONE
// kernel -- the two directions of the dual-stream residual exchange at one skip site (models/controlnet.py:1078-1087 and
// :2446-2461 read the same pair of co-located tensors): out0 = res0 + W0 a0 + b0, out1 = res1 + W1 a1 + b1.
int unib200_conv_gemm_dual(unib200_program* prog, const unib200_gemm_desc* d0, const unib200_gemm_desc* d1, void* stream) {
  PreparedGemm g0, g1;
  int rc = prepare_gemm(d0, &g0);
  if (rc != 0) return rc;
  if ((rc = prepare_gemm(d1, &g1)) != 0) return rc;
  if (d0->M != d1->M || d0->N != d1->N || d0->nseg != 1 || d1->nseg != 1 || d0->seg[0].C != d1->seg[0].C ||
      d0->seg[0].kind != UNIB200_SEG_1x1 || d1->seg[0].kind != UNIB200_SEG_1x1 || d0->flags != 0 || d1->flags != 0 ||
      d0->H != d1->H || d0->W != d1->W || d0->B != d1->B)
    return fail("conv_gemm_dual: the two problems must have the same shape, one 1x1 segment and no flags");
  if (g0.p.mode != 0 || g1.p.mode != 0 || g0.p.splits != 1 || g1.p.splits != 1 || g0.p.cg != g1.p.cg || g0.bn != g1.bn)
    return fail("conv_gemm_dual: both problems need the vector epilogue (aligned NHWC fp16 output, no split-K)");
  if (d0->bias_step || d1->bias_step || d0->bias_bstride || d1->bias_bstride || d0->rowstats_out || d1->rowstats_out ||
      d0->ln_rowstats || d1->ln_rowstats || !d0->bias != !d1->bias || !d0->res != !d1->res || d0->ldc != d1->ldc ||
      d0->ldr != d1->ldr || !d0->gn_part != !d1->gn_part || d0->gn_gran != d1->gn_gran || d0->gn_rows != d1->gn_rows)
    return fail("conv_gemm_dual: epilogue options of the two problems must match (plain bias / residual / GroupNorm "
                "statistics only)");
  GemmMaps maps = g0.maps;
  maps.a[1] = g1.maps.a[0];
  maps.b2 = g1.maps.b;
  GemmParams p = g0.p;
  p.dual = 1;
  p.mt_single = g0.p.m_tiles;
  p.m_tiles = 2 * g0.p.m_tiles;
  {
    auto recip = [](unsigned long long dv) { return (1ull << 40) / dv + 1ull; };
    if (static_cast<long long>(p.m_tiles) * p.n_tiles * 16 >= (1 << 20)) return fail("conv_gemm_dual: too many tiles");
    p.mul_tiles = recip(static_cast<unsigned long long>(p.m_tiles) * p.n_tiles);
  }
  p.bias2 = g1.p.bias;
  p.res2 = g1.p.res;
  p.out2 = g1.p.out;
  p.gn_part2 = g1.p.gn_part;
  const int sms = num_sms();
  const int bn = g0.bn;
  Op op = [maps, p, bn, sms](cudaStream_t s) { return launch_gemm(maps, p, bn, sms, s); };
  return submit(prog, std::move(op), 1, stream, "conv_gemm_dual", UNIB200_OP_GEMM, g0.flops + g1.flops, g0.bytes + g1.bytes,
                "dual " + g0.desc);
}

int unib200_conv_wgrad(unib200_program* prog, const unib200_wgrad_desc* d, void* stream) {
  if (!d || !d->x || !d->dy || !d->dw) return fail("conv_wgrad: null pointer");
  if (d->M <= 0 || d->N <= 0 || d->C <= 0 || (d->taps != 1 && d->taps != 9)) return fail("conv_wgrad: bad M / N / C / taps");
  if (d->ldx % 8 || d->lddy % 8 || d->ldx < d->C || d->lddy < d->N) return fail("conv_wgrad: leading dims must be multiples of 8");
  const bool linear = (d->H == 0 || d->W == 0);
  if (linear && d->taps != 1) return fail("conv_wgrad: plain matrices only support taps = 1");
  WgradMaps maps;
  memset(&maps, 0, sizeof(maps));
  WgradParams p;
  memset(&p, 0, sizeof(p));
  std::string why;
  int bw = 128, bh = 1, bb = 1;
  uint64_t Wd = d->M, Hd = 1, Bd = 1;
  if (linear) {
    p.w_shift = 30;
    p.h_shift = 0;
  } else {
    if (d->B * d->H * d->W != d->M) return fail("conv_wgrad: M != B*H*W");
    auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
    if (!pow2(d->W) || !pow2(d->H)) return fail("conv_wgrad: H and W must be powers of two");
    auto ilog2 = [](int x) { int s = 0; while ((1 << s) < x) ++s; return s; };
    p.w_shift = ilog2(d->W);
    p.h_shift = ilog2(d->H);
    bw = d->W < 128 ? d->W : 128;
    bh = 128 / bw;
    if (bh > d->H) bh = d->H;
    bb = 128 / (bw * bh);
    Wd = d->W; Hd = d->H; Bd = d->B;
  }
  auto mk = [&](CUtensorMap* m, const void* base, int Cn, int ld) {
    const uint64_t pix = static_cast<uint64_t>(ld) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(Cn), Wd, Hd, Bd};
    const uint64_t st[3] = {pix, pix * Wd, pix * Wd * Hd};
    const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bb)};
    return encode_map(m, base, 4, dims, st, box, &why);
  };
  if (!mk(&maps.dy, d->dy, d->N, d->lddy)) return fail("conv_wgrad dY map: " + why);
  if (!mk(&maps.x, d->x, d->C, d->ldx)) return fail("conv_wgrad X map: " + why);
  p.M = d->M; p.N = d->N; p.C = d->C; p.taps = d->taps;
  p.m_blocks = (d->M + 127) / 128;
  p.n_tiles = (d->N + 127) / 128;
  p.c_tiles = (d->C + wgrad_cin_tile() - 1) / wgrad_cin_tile();
  p.dw = d->dw;
  p.partial = d->partial;
  const int tiles = p.n_tiles * p.c_tiles * p.taps;
  int splits = (2 * num_sms() + tiles - 1) / tiles;
  if (splits > p.m_blocks) splits = p.m_blocks;
  const size_t slab = static_cast<size_t>(d->N) * d->taps * d->C * sizeof(float);
  if (!d->partial) splits = 1;
  else if (static_cast<size_t>(splits) * slab > d->partial_bytes) splits = static_cast<int>(d->partial_bytes / slab);
  if (splits < 1) splits = 1;
  p.splits = splits;
  const __half* dy = static_cast<const __half*>(d->dy);
  float* db = d->db;
  const int lddy = d->lddy, M = d->M, N = d->N;
  if (db && N % 8) return fail("conv_wgrad: the bias gradient needs N to be a multiple of 8");
  Op op = [maps, p, dy, db, lddy, M, N](cudaStream_t s) {
    cudaError_t e = launch_wgrad(maps, p, s);
    if (e == cudaSuccess && db) e = launch_colsum(dy, lddy, M, N, db, s);
    return e;
  };
  const double flops = 2.0 * d->M * d->N * d->C * d->taps;
  return submit(prog, std::move(op), (splits > 1 ? 2 : 1) + (db ? 1 : 0), stream, "conv_wgrad", UNIB200_OP_GEMM, flops,
                2.0 * d->M * (d->N + d->C) + 4.0 * d->N * d->taps * d->C,
                "wgrad M=" + std::to_string(d->M) + " N=" + std::to_string(d->N) + " C=" + std::to_string(d->C) + " taps=" +
                    std::to_string(d->taps) + " splits=" + std::to_string(splits));
}

int unib200_pack_master_weight(unib200_program* prog, const float* w, int O, int I, int taps, void* out, int dgrad,
                               void* stream) {
  if (!w || !out || O <= 0 || I <= 0 || (taps != 1 && taps != 9)) return fail("pack_master_weight: bad arguments (taps 1 or 9)");
  __half* op_ = static_cast<__half*>(out);
  Op op = [=](cudaStream_t s) { return launch_pack_master(w, op_, O, I, taps, dgrad, s); };
  return submit(prog, std::move(op), 1, stream, "pack_master_weight");
}

int unib200_wgrad_scatter_add(unib200_program* prog, const float* dw, float* grad, int N, int taps, int C, void* stream) {
  if (!dw || !grad || N <= 0 || C <= 0 || taps <= 0) return fail("wgrad_scatter_add: bad arguments");
  Op op = [=](cudaStream_t s) { return launch_wgrad_scatter_add(dw, grad, N, taps, C, s); };
  return submit(prog, std::move(op), 1, stream, "wgrad_scatter_add");
}

int unib200_colsum(unib200_program* prog, const void* x, int ld, int M, int N, float* out, void* stream) {
  if (!x || !out || M <= 0 || N <= 0 || N % 8 || ld % 8 || ld < N) return fail("colsum: bad arguments (N and ld multiples of 8)");
  const __half* xp = static_cast<const __half*>(x);
  Op op = [=](cudaStream_t s) { return launch_colsum(xp, ld, M, N, out, s); };
  return submit(prog, std::move(op), 1, stream, "colsum");
}

int unib200_layernorm_backward(unib200_program* prog, const void* x, const void* dy, void* dx, const float* gamma,
                               float* dgamma_dbeta, float* scratch, size_t scratch_floats, int rows, int C, float eps,
                               void* stream) {
  if (!x || !dy || !dx || !gamma || !dgamma_dbeta || !scratch || rows <= 0 || C <= 0 || C > 1536)
    return fail("layernorm_backward: bad arguments (C <= 1536)");
  const int max_slabs = static_cast<int>(scratch_floats / (2 * static_cast<size_t>(C)));
  if (max_slabs < 1) return fail("layernorm_backward: scratch too small");
  const __half *xp = static_cast<const __half*>(x), *dp = static_cast<const __half*>(dy);
  __half* dxp = static_cast<__half*>(dx);
  Op op = [=](cudaStream_t s) {
    return launch_layernorm_backward(xp, dp, dxp, gamma, dgamma_dbeta, nullptr, scratch, max_slabs, rows, C, eps, s);
  };
  return submit(prog, std::move(op), 2, stream, "layernorm_backward", UNIB200_OP_LAYERNORM, 0.0, 6.0 * rows * C);
}

int unib200_geglu(unib200_program* prog, const void* proj, const void* dout, void* out, int64_t rows, int inner, void* stream) {
  if (!proj || !out || rows <= 0 || inner <= 0) return fail("geglu: bad arguments");
  const __half *pp = static_cast<const __half*>(proj), *dp = static_cast<const __half*>(dout);
  __half* op_ = static_cast<__half*>(out);
  Op op = [=](cudaStream_t s) { return launch_geglu(pp, dp, op_, rows, inner, dp != nullptr, s); };
  return submit(prog, std::move(op), 1, stream, "geglu");
}

int unib200_softmax_backward(unib200_program* prog, const void* P, void* dP, int rows, int n, int ld, float scale, void* stream) {
  if (!P || !dP || rows <= 0 || n <= 0 || ld < n) return fail("softmax_backward: bad arguments");
  const __half* pp = static_cast<const __half*>(P);
  __half* dp = static_cast<__half*>(dP);
  Op op = [=](cudaStream_t s) { return launch_softmax_backward(pp, dp, rows, n, ld, scale, s); };
  return submit(prog, std::move(op), 1, stream, "softmax_backward");
}

int unib200_cvt_f32_f16(unib200_program* prog, const float* src, void* dst, int64_t rows, int cols, int ld, void* stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || ld < cols) return fail("cvt_f32_f16: bad arguments");
  __half* dp = static_cast<__half*>(dst);
  Op op = [=](cudaStream_t s) { return launch_cvt_f32_f16(src, dp, rows, cols, ld, s); };
  return submit(prog, std::move(op), 1, stream, "cvt_f32_f16");
}

int unib200_silu_f16(unib200_program* prog, const void* x, const void* dy, void* out, int64_t n, void* stream) {
  if (!x || !out || n <= 0) return fail("silu_f16: bad arguments");
  const __half *xp = static_cast<const __half*>(x), *dp = static_cast<const __half*>(dy);
  __half* op_ = static_cast<__half*>(out);
  Op op = [=](cudaStream_t s) { return launch_silu_f16(xp, dp, op_, n, s); };
  return submit(prog, std::move(op), 1, stream, "silu_f16");
}

int unib200_pool2x2_sum(unib200_program* prog, const void* src, void* dst, int B, int H, int W, int C, void* stream) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return fail("pool2x2_sum: bad arguments (C a multiple of 8)");
  const __half* sp = static_cast<const __half*>(src);
  __half* dp = static_cast<__half*>(dst);
  Op op = [=](cudaStream_t s) { return launch_pool2x2_sum(sp, dp, B, H, W, C, s); };
  return submit(prog, std::move(op), 1, stream, "pool2x2_sum");
}

int unib200_scatter2x(unib200_program* prog, const void* src, void* dst, int B, int H, int W, int C, void* stream) {
  if (!src || !dst || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return fail("scatter2x: bad arguments (C a multiple of 8)");
  const __half* sp = static_cast<const __half*>(src);
  __half* dp = static_cast<__half*>(dst);
  Op op = [=](cudaStream_t s) { return launch_scatter2x(sp, dp, B, H, W, C, s); };
  return submit(prog, std::move(op), 1, stream, "scatter2x");
}

int unib200_adamw_step(unib200_program* prog, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) return fail("adamw_step: bad arguments (step counts from 1)");
  Op op = [=](cudaStream_t s) { return launch_adamw(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, s); };
  return submit(prog, std::move(op), 1, stream, "adamw_step");
}

int unib200_groupnorm_backward(unib200_program* prog, const unib200_gn_bwd_desc* d, void* stream) {
  if (!d || !d->x || !d->dz || !d->dx || !d->gamma || !d->beta || !d->dgamma || !d->dbeta || !d->scratch)
    return fail("groupnorm_backward: null pointer");
  if (d->groups <= 0 || d->C % d->groups || d->C / d->groups > 256 || d->B <= 0 || d->HW <= 0)
    return fail("groupnorm_backward: unsupported channel / group configuration");
  GnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.x = static_cast<const __half*>(d->x); p.ldx = d->ldx;
  p.dz = static_cast<const __half*>(d->dz); p.ldz = d->ldz;
  p.dx = static_cast<__half*>(d->dx); p.lddx = d->lddx;
  p.gamma = d->gamma; p.beta = d->beta;
  p.dgamma_part = d->scratch;
  p.dbeta_part = d->scratch + static_cast<size_t>(d->B) * d->C;
  p.HW = d->HW; p.C = d->C; p.G = d->groups; p.silu = d->silu; p.eps = d->eps;
  const int B = d->B, Cn = d->C;
  float *dg = d->dgamma, *dbt = d->dbeta;
  Op op = [p, B, Cn, dg, dbt](cudaStream_t s) {
    cudaError_t e = launch_gn_backward(p, B, s);
    if (e == cudaSuccess) e = launch_sum_slabs(p.dgamma_part, dg, Cn, B, s);
    if (e == cudaSuccess) e = launch_sum_slabs(p.dbeta_part, dbt, Cn, B, s);
    return e;
  };
  return submit(prog, std::move(op), 3, stream, "groupnorm_backward", UNIB200_OP_GROUPNORM, 0.0,
                6.0 * static_cast<double>(d->B) * d->HW * d->C);
}

int unib200_attention(unib200_program* prog, const unib200_attn_desc* d, void* stream) {
  if (!d) return fail("null desc");
  if (d->d % 8 != 0 || d->d < 8 || d->d > 192) return fail("attention: head dim must be a multiple of 8 in [8,192]");
  if (d->ldq % 8 || d->ldk % 8 || d->ldv % 8 || d->ldo % 8) return fail("attention: leading dims must be multiples of 8");
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string why;
  const int bkv = attention_bkv(d->d);
  auto mk = [&](CUtensorMap* m, const void* base, int ld, int tokens, int box_rows) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->d), static_cast<uint64_t>(tokens),
                              static_cast<uint64_t>(d->heads), static_cast<uint64_t>(d->B)};
    const uint64_t st[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(d->d) * 2,
                            static_cast<uint64_t>(ld) * 2 * tokens};
    const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), 1, 1};
    return encode_map(m, base, 4, dims, st, box, &why);
  };
  if (!mk(&maps.q, d->q, d->ldq, d->Nq, 128)) return fail("attention Q map: " + why);
  if (!mk(&maps.k, d->k, d->ldk, d->Nk, bkv)) return fail("attention K map: " + why);
  if (!mk(&maps.v, d->v, d->ldv, d->Nk, bkv)) return fail("attention V map: " + why);
  AttnParams p;
  p.B = d->B; p.heads = d->heads; p.Nq = d->Nq; p.Nk = d->Nk; p.d = d->d;
  p.scale = d->scale;
  p.out = static_cast<__half*>(d->out);
  p.ldo = d->ldo;
  p.lse2 = d->lse2;
  p.trace = g_attn_trace;
  p.pdl_early = unib::g_pdl_enabled == 2 ? 1 : 0;
  Op op = [maps, p](cudaStream_t s) { return launch_attention(maps, p, s); };
  // what-if timing aid (garbage results): UNIB200_SKIP_ATTN=1 drops the short launches (cross-attention, <= 1024 tokens),
  // 2 the long self-attention launches
  static const int skip_attn = getenv("UNIB200_SKIP_ATTN") ? atoi(getenv("UNIB200_SKIP_ATTN")) : 0;
  if (prog && skip_attn) {
    const bool small = d->Nk < 512 || d->Nq <= 1024;
    if ((skip_attn & 1) && small) op = [](cudaStream_t) { return cudaSuccess; };
    if ((skip_attn & 2) && !small) op = [](cudaStream_t) { return cudaSuccess; };
  }
  const double bh = static_cast<double>(d->B) * d->heads;
  return submit(prog, std::move(op), 1, stream, "attention", UNIB200_OP_ATTENTION, 4.0 * bh * d->Nq * d->Nk * d->d,
                2.0 * bh * d->d * (2.0 * d->Nq + 2.0 * d->Nk),
                "attention B=" + std::to_string(d->B) + " h=" + std::to_string(d->heads) + " Nq=" +
                    std::to_string(d->Nq) + " Nk=" + std::to_string(d->Nk) + " d=" + std::to_string(d->d));
}

int unib200_attention_backward(unib200_program* prog, const unib200_attn_bwd_desc* d, void* stream) {
  if (!d || !d->q || !d->k || !d->v || !d->o || !d->dout || !d->lse2 || !d->D || !d->dq_acc || !d->dk || !d->dv)
    return fail("attention_backward: null pointer");
  if (d->d % 8 != 0 || d->d < 8 || d->d > 80) return fail("attention_backward: head dim must be a multiple of 8 in [8,80]");
  if (d->ldq % 8 || d->ldk % 8 || d->ldv % 8 || d->ldo % 8 || d->lddo % 8 || d->ld_dk % 8 || d->ld_dv % 8)
    return fail("attention_backward: leading dims must be multiples of 8");
  AttnBwdMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string why;
  auto mk = [&](CUtensorMap* m, const void* base, int ld, int tokens) {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->d), static_cast<uint64_t>(tokens),
                              static_cast<uint64_t>(d->heads), static_cast<uint64_t>(d->B)};
    const uint64_t st[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(d->d) * 2,
                            static_cast<uint64_t>(ld) * 2 * tokens};
    const uint32_t box[4] = {64, 128, 1, 1};
    return encode_map(m, base, 4, dims, st, box, &why);
  };
  if (!mk(&maps.q, d->q, d->ldq, d->Nq)) return fail("attention_backward Q map: " + why);
  if (!mk(&maps.k, d->k, d->ldk, d->Nk)) return fail("attention_backward K map: " + why);
  if (!mk(&maps.v, d->v, d->ldv, d->Nk)) return fail("attention_backward V map: " + why);
  if (!mk(&maps.dout, d->dout, d->lddo, d->Nq)) return fail("attention_backward dO map: " + why);
  AttnBwdParams p;
  p.B = d->B; p.heads = d->heads; p.Nq = d->Nq; p.Nk = d->Nk; p.d = d->d;
  p.scale = d->scale;
  p.lse2 = d->lse2; p.D = d->D; p.dq_acc = d->dq_acc; p.ld_dq = d->ld_dq;
  p.dk = static_cast<__half*>(d->dk); p.ld_dk = d->ld_dk;
  p.dv = static_cast<__half*>(d->dv); p.ld_dv = d->ld_dv;
  const __half* o = static_cast<const __half*>(d->o);
  const __half* dout = static_cast<const __half*>(d->dout);
  const int ldo = d->ldo, lddo = d->lddo;
  Op op = [maps, p, o, dout, ldo, lddo](cudaStream_t s) { return launch_attention_bwd(maps, p, o, ldo, dout, lddo, s); };
  const double bh = static_cast<double>(d->B) * d->heads;
  return submit(prog, std::move(op), 2, stream, "attention_backward", UNIB200_OP_ATTENTION,
                10.0 * bh * d->Nq * d->Nk * d->d, 2.0 * bh * d->d * (4.0 * d->Nq + 4.0 * d->Nk));
}

int unib200_groupnorm(unib200_program* prog, const unib200_gn_desc* d, void* stream) {
  if (!d) return fail("null desc");
  GnParams p;
  memset(&p, 0, sizeof(p));
  p.x1 = static_cast<const __half*>(d->x1); p.ld1 = d->ld1; p.C1 = d->C1;
  p.x2 = static_cast<const __half*>(d->x2); p.ld2 = d->ld2; p.C2 = d->x2 ? d->C2 : 0;
  p.HW = d->HW; p.G = d->groups; p.eps = d->eps; p.gamma = d->gamma; p.beta = d->beta;
  p.out = static_cast<__half*>(d->out); p.silu = d->silu; p.partial = d->scratch;
  const int C = p.C1 + p.C2;
  if (C % 8 || p.C1 % 8 || d->groups <= 0 || C % d->groups || C / 8 > 512 || d->ld1 % 8 ||
      (d->x2 && d->ld2 % 8))
    return fail("groupnorm: unsupported channel configuration");
  const int B = d->B, sms = num_sms();
  if (d->part1) {
    // statistics come from the producing GEMMs' epilogues: one apply launch
    const int g = d->part_gran, R = d->part_rows;
    if (d->x2 && !d->part2) return fail("groupnorm: part2 is required when x2 and part1 are given");
    if (g < 2 || p.C1 % g || p.C2 % g || (C / d->groups) % g || R < 1 || d->HW % R)
      return fail("groupnorm: part_gran must divide C1, C2 and the group size; part_rows must divide HW");
    if (C / g > 1024) return fail("groupnorm: too many micro-groups");
    p.part1 = d->part1; p.part2 = d->part2; p.part_gran = g; p.part_rows = R;
    Op op = [p, B, sms](cudaStream_t s) { return launch_groupnorm_parts(p, B, sms, s); };
    return submit(prog, std::move(op), 1, stream, "groupnorm", UNIB200_OP_GROUPNORM, 0.0,
                  4.0 * static_cast<double>(d->B) * d->HW * C,
                  "groupnorm(fused stats) B=" + std::to_string(d->B) + " HW=" + std::to_string(d->HW) + " C=" +
                      std::to_string(p.C1) + "+" + std::to_string(p.C2));
  }
  const size_t per_chunk = static_cast<size_t>(d->B) * d->groups * 2;
  const size_t mc = per_chunk ? d->scratch_floats / per_chunk : 0;
  if (mc < 1 || !d->scratch) return fail("groupnorm: scratch too small");
  p.max_chunks = mc > 4096 ? 4096 : static_cast<int>(mc);
  Op op = [p, B, sms](cudaStream_t s) { return launch_groupnorm(p, B, sms, s); };
  return submit(prog, std::move(op), d->HW <= 256 ? 1 : 2, stream, "groupnorm", UNIB200_OP_GROUPNORM, 0.0,
                4.0 * static_cast<double>(d->B) * d->HW * C,
                "groupnorm B=" + std::to_string(d->B) + " HW=" + std::to_string(d->HW) + " C=" + std::to_string(p.C1) +
                    "+" + std::to_string(p.C2));
}

int unib200_layernorm(unib200_program* prog, const void* x, void* y, const float* gamma, const float* beta, int rows,
                      int C, float eps, void* stream) {
  if (C % 8 || C > 2048) return fail("layernorm: C must be a multiple of 8 and <= 2048");
  Op op = [=](cudaStream_t s) {
    return launch_layernorm(static_cast<const __half*>(x), static_cast<__half*>(y), gamma, beta, rows, C, eps, s);
  };
  return submit(prog, std::move(op), 1, stream, "layernorm", UNIB200_OP_LAYERNORM, 0.0, 4.0 * rows * C,
                "layernorm rows=" + std::to_string(rows) + " C=" + std::to_string(C));
}

int unib200_to_nhwc(unib200_program* prog, const void* src, int src_is_f32, void* dst, int B, int C, int H, int W,
                    int64_t sb, int64_t sc, int64_t sh, int64_t sw, int Cpad, void* stream) {
  if (Cpad < C) return fail("to_nhwc: Cpad < C");
  Op op = [=](cudaStream_t s) {
    return launch_to_nhwc(src, src_is_f32, static_cast<__half*>(dst), B, C, H, W, sb, sc, sh, sw, Cpad, s);
  };
  return submit(prog, std::move(op), 1, stream, "to_nhwc");
}

int unib200_from_nhwc(unib200_program* prog, const void* src, void* dst, int dst_is_f32, int B, int C, int HW, int ld,
                      void* stream) {
  Op op = [=](cudaStream_t s) {
    return launch_from_nhwc(static_cast<const __half*>(src), dst, dst_is_f32, B, C, HW, ld, s);
  };
  return submit(prog, std::move(op), 1, stream, "from_nhwc");
}

int unib200_upsample2x(unib200_program* prog, const void* src, void* dst, int B, int H, int W, int C, void* stream) {
  if (C % 8) return fail("upsample2x: C must be a multiple of 8");
  Op op = [=](cudaStream_t s) {
    return launch_upsample2x(static_cast<const __half*>(src), static_cast<__half*>(dst), B, H, W, C, s);
  };
  return submit(prog, std::move(op), 1, stream, "upsample2x");
}

int unib200_timestep_sinusoid(unib200_program* prog, const float* t, const int* step_idx, int t_stride, float* out,
                              int B, int dim, void* stream) {
  if (dim % 2) return fail("timestep_sinusoid: dim must be even");
  Op op = [=](cudaStream_t s) { return launch_timestep_sinusoid(t, step_idx, t_stride, out, B, dim, s); };
  return submit(prog, std::move(op), 1, stream, "timestep_sinusoid");
}

int unib200_gemv(unib200_program* prog, const float* x, const void* w_fp16, const float* bias, float* y, int B, int K,
                 int N, int act_silu, void* stream) {
  if (K % 8 || static_cast<size_t>(K) * 8 * 4 > 48 * 1024) return fail("gemv: K must be a multiple of 8 and <= 1536");
  Op op = [=](cudaStream_t s) {
    return launch_gemv(x, static_cast<const __half*>(w_fp16), bias, y, B, K, N, act_silu, s);
  };
  return submit(prog, std::move(op), (B + 7) / 8, stream, "gemv");
}

int unib200_axpby(unib200_program* prog, const float* model_out, const float* x, float* out, const float* coef,
                  const int* step_idx, int64_t n, void* stream) {
  Op op = [=](cudaStream_t s) { return launch_axpby(model_out, x, out, coef, step_idx, n, s); };
  return submit(prog, std::move(op), 1, stream, "axpby");
}

int unib200_unipc_step(unib200_program* prog, const float* model_out, float* sample, float* last_sample, float* hist0,
                       float* hist1, const float* coef, const int* step_idx, int B, int C, int HW, int first_channel,
                       void* stream) {
  if (B <= 0 || C <= 0 || HW <= 0 || first_channel < 0 || first_channel >= C) return fail("unipc_step: bad shape");
  if (!model_out || !sample || !last_sample || !hist0 || !hist1 || !coef) return fail("unipc_step: null pointer");
  Op op = [=](cudaStream_t s) {
    return launch_unipc_step(model_out, sample, last_sample, hist0, hist1, coef, step_idx, B, C, HW, first_channel, s);
  };
  return submit(prog, std::move(op), 1, stream, "unipc_step");
}

int unib200_add_f16(unib200_program* prog, const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (n % 8) return fail("add_f16: n must be a multiple of 8");
  Op op = [=](cudaStream_t s) {
    return launch_add_f16(static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<__half*>(out), n, s);
  };
  return submit(prog, std::move(op), 1, stream, "add_f16");
}

int unib200_softmax_rows(unib200_program* prog, void* s, int rows, int n, int ld, float scale, void* stream) {
  if (rows <= 0 || n <= 0 || ld % 8 || ld < (n + 7) / 8 * 8)
    return fail("softmax_rows: ld must be a multiple of 8 and cover n rounded up to 8 (columns n .. are written as zeros)");
  if (!(scale > 0.f)) return fail("softmax_rows: scale must be positive");
  if (reinterpret_cast<uintptr_t>(s) & 15) return fail("softmax_rows: matrix must be 16-byte aligned");
  Op op = [=](cudaStream_t st) { return launch_softmax_rows(static_cast<__half*>(s), rows, n, ld, scale, st); };
  return submit(prog, std::move(op), 1, stream, "softmax_rows");
}

int unib200_gaussian_sample(unib200_program* prog, const float* moments, const float* noise, float* out, int B, int C,
                            int HW, float scale, void* stream) {
  if (B <= 0 || C <= 0 || HW <= 0 || !moments || !out) return fail("gaussian_sample: bad arguments");
  Op op = [=](cudaStream_t st) { return launch_gaussian_sample(moments, noise, out, B, C, HW, scale, st); };
  return submit(prog, std::move(op), 1, stream, "gaussian_sample");
}

// ---------------------------------------------------------------------------------------------------------------
// step-level context
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

#include <map>

struct unib200_ctx {
  struct Buf { void* ptr = nullptr; size_t bytes = 0; bool owned = false; };
  int device = 0;
  int use_graph = 1;
  std::map<std::string, Buf> bufs;
  std::map<std::string, unib200_program*> progs;
};

namespace {
unib200_ctx::Buf* ctx_buf(unib200_ctx* c, const char* key) {
  auto it = c->bufs.find(key);
  return it == c->bufs.end() ? nullptr : &it->second;
}
int ctx_put(unib200_ctx* c, const char* key, void* ptr, size_t bytes, bool owned) {
  unib200_ctx::Buf& b = c->bufs[key];
  if (b.owned && b.ptr) cudaFree(b.ptr);
  b.ptr = ptr; b.bytes = bytes; b.owned = owned;
  return 0;
}
int ctx_replay(unib200_ctx* c, const char* name, void* stream, bool allow_graph) {
  auto it = c->progs.find(name);
  if (it == c->progs.end()) return fail(std::string("context has no program named '") + name + "'");
  unib200_program* p = it->second;
  if (allow_graph && c->use_graph && p->exec) return unib200_program_graph_launch(p, stream);
  return unib200_program_run(p, stream);
}
}  // namespace

extern "C" {

unib200_ctx* unib200_create(int device, const unib200_config* cfg) {
  if (cudaSetDevice(device) != cudaSuccess) { fail("unib200_create: cudaSetDevice failed"); return nullptr; }
  unib200_ctx* c = new (std::nothrow) unib200_ctx();
  if (!c) return nullptr;
  c->device = device;
  c->use_graph = cfg ? cfg->use_graph : 1;
  return c;
}

void unib200_destroy(unib200_ctx* c) {
  if (!c) return;
  for (auto& kv : c->progs) unib200_program_destroy(kv.second);
  for (auto& kv : c->bufs)
    if (kv.second.owned && kv.second.ptr) cudaFree(kv.second.ptr);
  delete c;
}

int unib200_alloc(unib200_ctx* c, const char* key, size_t bytes, int zero) {
  if (!c || !key || bytes == 0) return fail("unib200_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail_cuda("unib200_alloc", e);
  if (zero && (e = cudaMemset(p, 0, bytes)) != cudaSuccess) { cudaFree(p); return fail_cuda("unib200_alloc (memset)", e); }
  return ctx_put(c, key, p, bytes, true);
}

int unib200_bind(unib200_ctx* c, const char* key, void* dev_ptr, size_t bytes) {
  if (!c || !key || !dev_ptr) return fail("unib200_bind: bad arguments");
  return ctx_put(c, key, dev_ptr, bytes, false);
}

void* unib200_buffer(unib200_ctx* c, const char* key, size_t* bytes) {
  unib200_ctx::Buf* b = (c && key) ? ctx_buf(c, key) : nullptr;
  if (bytes) *bytes = b ? b->bytes : 0;
  return b ? b->ptr : nullptr;
}

int unib200_load_weight(unib200_ctx* c, const char* key, const void* src, int dtype, const int64_t* shape, int ndim) {
  if (!c || !key || !src || !shape || ndim < 1 || ndim > 8) return fail("unib200_load_weight: bad arguments");
  const size_t es = dtype == UNIB200_F16 ? 2 : (dtype == UNIB200_F32 || dtype == UNIB200_I32) ? 4 : 0;
  if (!es) return fail("unib200_load_weight: dtype must be UNIB200_F16 / _F32 / _I32");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] <= 0) return fail("unib200_load_weight: non-positive dimension");
    n *= static_cast<size_t>(shape[i]);
  }
  int rc = unib200_alloc(c, key, n * es, 0);
  if (rc != 0) return rc;
  cudaError_t e = cudaMemcpy(ctx_buf(c, key)->ptr, src, n * es, cudaMemcpyDefault);     // host or device source
  if (e != cudaSuccess) return fail_cuda("unib200_load_weight (copy)", e);
  return 0;
}

int unib200_ctx_attach(unib200_ctx* c, const char* name, unib200_program* prog) {
  if (!c || !name || !prog) return fail("unib200_ctx_attach: bad arguments");
  auto it = c->progs.find(name);
  if (it != c->progs.end() && it->second != prog) unib200_program_destroy(it->second);
  c->progs[name] = prog;
  return 0;
}

int unib200_ctx_run(unib200_ctx* c, const char* name, void* stream) {
  if (!c || !name) return fail("unib200_ctx_run: bad arguments");
  return ctx_replay(c, name, stream, true);
}
int unib200_unet_forward(unib200_ctx* c, void* stream) { return unib200_ctx_run(c, "unet", stream); }
int unib200_attr_enc_forward(unib200_ctx* c, void* stream) { return unib200_ctx_run(c, "attr_enc", stream); }
int unib200_attr_dec_forward(unib200_ctx* c, void* stream) { return unib200_ctx_run(c, "attr_dec", stream); }
int unib200_dual_step(unib200_ctx* c, void* stream) { return unib200_ctx_run(c, "step", stream); }

int unib200_sample_loop(unib200_ctx* c, int n_steps, const void* lat_img, const void* lat_attr, const void* ehs,
                        void* out_lat_img, void* out_lat_attr, void* stream) {
  if (!c || n_steps < 0) return fail("unib200_sample_loop: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unib200_ctx::Buf *bi = ctx_buf(c, "lat_img"), *ba = ctx_buf(c, "lat_attr"), *be = ctx_buf(c, "ehs"),
                   *bs = ctx_buf(c, "step");
  if (!bi || !ba || !be || !bs) return fail("unib200_sample_loop: bind lat_img, lat_attr, ehs and step first");
  cudaError_t e;
  if (lat_img && (e = cudaMemcpyAsync(bi->ptr, lat_img, bi->bytes, cudaMemcpyDefault, s)) != cudaSuccess)
    return fail_cuda("sample_loop (lat_img in)", e);
  if (lat_attr && (e = cudaMemcpyAsync(ba->ptr, lat_attr, ba->bytes, cudaMemcpyDefault, s)) != cudaSuccess)
    return fail_cuda("sample_loop (lat_attr in)", e);
  if (ehs && (e = cudaMemcpyAsync(be->ptr, ehs, be->bytes, cudaMemcpyDefault, s)) != cudaSuccess)
    return fail_cuda("sample_loop (ehs in)", e);
  if ((e = cudaMemsetAsync(bs->ptr, 0, bs->bytes, s)) != cudaSuccess) return fail_cuda("sample_loop (step counter)", e);
  int rc = 0;
  if (c->progs.count("setup") && (rc = ctx_replay(c, "setup", stream, false)) != 0) return rc;
  for (int i = 0; i < n_steps; ++i)
    if ((rc = ctx_replay(c, "step", stream, true)) != 0) return rc;
  if (out_lat_img && (e = cudaMemcpyAsync(out_lat_img, bi->ptr, bi->bytes, cudaMemcpyDefault, s)) != cudaSuccess)
    return fail_cuda("sample_loop (lat_img out)", e);
  if (out_lat_attr && (e = cudaMemcpyAsync(out_lat_attr, ba->ptr, ba->bytes, cudaMemcpyDefault, s)) != cudaSuccess)
    return fail_cuda("sample_loop (lat_attr out)", e);
  return 0;
}

int unib200_add_int(unib200_program* prog, int* p, int v, void* stream) {
  Op op = [=](cudaStream_t s) { return launch_add_int(p, v, s); };
  return submit(prog, std::move(op), 1, stream, "add_int");
}

}  // extern "C"
