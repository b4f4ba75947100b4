#!/usr/bin/env python
"""In-kernel phase timeline of the GEMM kernel (globaltimer stamps of CTA 0, unib200_debug_set_trace) for the shapes of
tools/bench_gemm.py, replayed as a dependent chain inside a CUDA graph.  Columns are ns relative to kernel entry:
setup = barriers/TMEM ready, tma0 = first TMA issued, mma0 = first stage landed, acc0 = first tile's MMAs committed,
epi0 = epilogue saw the first accumulator, drain = last TMA store issued, done = stores complete, exit; gap = entry
minus the previous launch's exit (launch latency inside the graph).  Debug instrument."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the stamps are compiled in only with -DUNIB_GEMM_TRACE: build that variant here (`python -m uni_renderer_b200.build
# --trace`, no GPU needed) and it is picked up through UNIB200_LIB
_TRACE_LIB = os.path.join(ROOT, "uni_renderer_b200", "libunib200_trace.so")
if not os.path.exists(_TRACE_LIB):
    raise SystemExit("build the trace variant first: python -m uni_renderer_b200.build --trace")
os.environ["UNIB200_LIB"] = _TRACE_LIB
import torch  # noqa: E402

from uni_renderer_b200 import _lib, ops  # noqa: E402
from uni_renderer_b200.ops import EPI_GEGLU, SEG_1x1, SEG_3x3_S2  # noqa: E402
from tools.bench_gemm import SHAPES  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=6)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
partial = torch.empty(16 << 20, device=dev, dtype=torch.float32)
lib = _lib.load()
names = ["setup", "tma0", "mma0", "acc0", "epi0", "drain", "done", "exit"]
print(f"{'shape':26s} {'gap':>6s} " + " ".join(f"{n:>6s}" for n in names))
for i, (name, B, H, cins, kind, cout, res, geglu) in enumerate(SHAPES):
    M = B * H * H
    Hin = 2 * H if kind == SEG_3x3_S2 else H
    srcs = [torch.randn(B * Hin * Hin, c, device=dev).half() for c in cins]
    w = ops.pack_weight([(torch.randn(cout, c, *((1, 1) if kind == SEG_1x1 else (3, 3)), device=dev) * 0.02, kind)
                         for c in cins])
    bias = torch.randn(cout, device=dev)
    flags, n_out = (EPI_GEGLU, cout // 2) if geglu else (0, cout)
    outs = [torch.empty(M, n_out, device=dev, dtype=torch.float16) for _ in range(2)]
    r = torch.randn(M, n_out, device=dev).half() if res else None
    tr = torch.zeros(16 + 16 * 4096, device=dev, dtype=torch.int64)
    lib.unib200_debug_set_trace(tr.data_ptr())
    prog = ops.Program()
    for k in range(a.reps):
        ops.conv_gemm(prog, [(s, c, kind) for s, c in zip(srcs, cins)], w, outs[k & 1], M=M, N=cout, B=B,
                      H=0 if kind == SEG_1x1 else H, W=0 if kind == SEG_1x1 else H, bias=bias, res=r, flags=flags,
                      partial=None if geglu else partial)
    lib.unib200_debug_set_trace(None)
    prog.run()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        prog.instantiate_graph()
        side.synchronize()
        for _ in range(3):
            prog.launch_graph()
        side.synchronize()
        tr.zero_()
        prog.launch_graph()
        side.synchronize()
    t = tr.cpu()
    n = int(t[0])
    rows = t[16:16 + 16 * n].reshape(n, 16)
    k = n - 1                      # last launch of the chain (steady state)
    e = rows[k]
    gap = int(e[0] - rows[k - 1][8]) if k > 0 else 0
    rel = [int(e[j] - e[0]) if e[j] else -1 for j in range(1, 9)]
    ext = [int(e[j] - e[0]) if e[j] else -1 for j in (14, 15, 9, 11)]
    sub = [int(e[j] - e[0]) if e[j] else -1 for j in (4, 10, 12, 13)]
    print(f"{name:26s} {gap:6d} " + " ".join(f"{v:6d}" for v in rel) + "   pdlwait/decoded/epi0end/epi1: " +
          " ".join(f"{v:6d}" for v in ext) + "   sub-tile 1 top/ld/math/stored: " + " ".join(f"{v:6d}" for v in sub))
