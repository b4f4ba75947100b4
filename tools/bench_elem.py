#!/usr/bin/env python
"""In-graph per-launch time of the GroupNorm / LayerNorm shapes of the step (chain of `reps` launches replayed as a
CUDA graph, CUDA events).  UNIB200_GN_TWO_KERNEL=1 selects the stats + apply pair.  Optimisation instrument."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uni_renderer_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
GN = [(4, 4096, 320, 0), (4, 1024, 640, 0), (4, 256, 1280, 0), (4, 64, 1280, 0), (4, 4096, 320, 320), (4, 4096, 640, 320),
      (4, 64, 1280, 1280), (4, 1024, 1280, 640)]
LN = [(16384, 320), (4096, 640), (1024, 1280), (256, 1280)]
reps, iters = 20, 10


def timed(prog):
    prog.run()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(side):
        prog.instantiate_graph()
        side.synchronize()
        for _ in range(3):
            prog.launch_graph()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            prog.launch_graph()
        e1.record()
        side.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)


scratch = torch.empty(1 << 18, device=dev, dtype=torch.float32)
for (B, HW, C1, C2) in GN:
    x1 = torch.randn(B * HW, C1, device=dev).half()
    x2 = torch.randn(B * HW, C2, device=dev).half() if C2 else None
    C = C1 + C2
    g, bt = torch.randn(C, device=dev), torch.randn(C, device=dev)
    outs = [torch.empty(B * HW, C, device=dev, dtype=torch.float16) for _ in range(2)]
    prog = ops.Program()
    for k in range(reps):
        ops.groupnorm(prog, x1, C1, x2, C2, g, bt, outs[k & 1], scratch, B=B, HW=HW, groups=32, eps=1e-5, silu=True)
    us = timed(prog)
    by = 4.0 * B * HW * C
    print(f"groupnorm B={B} HW={HW} C={C1}+{C2}: {us:7.2f} us  {by / us / 1e3:7.1f} GB/s")
# the same shapes with the statistics already in (row block, micro-group) partials: ONE apply launch (gn_apply_parts)
for (B, HW, C1, C2) in GN:
    if HW < 1024:
        continue
    gran, rows = 10, 128
    x1 = torch.randn(B * HW, C1, device=dev).half()
    x2 = torch.randn(B * HW, C2, device=dev).half() if C2 else None
    C = C1 + C2
    p1 = torch.rand(B * HW // rows, C1 // gran, 2, device=dev) * rows * gran
    p2 = torch.rand(B * HW // rows, C2 // gran, 2, device=dev) * rows * gran if C2 else None
    g, bt = torch.randn(C, device=dev), torch.randn(C, device=dev)
    outs = [torch.empty(B * HW, C, device=dev, dtype=torch.float16) for _ in range(2)]
    prog = ops.Program()
    for k in range(reps):
        ops.groupnorm(prog, x1, C1, x2, C2, g, bt, outs[k & 1], scratch, B=B, HW=HW, groups=32, eps=1e-5, silu=True,
                      parts=(p1, p2, gran, rows))
    us = timed(prog)
    print(f"groupnorm(parts) B={B} HW={HW} C={C1}+{C2}: {us:7.2f} us  {4.0 * B * HW * C / us / 1e3:7.1f} GB/s")
for (rows, C) in LN:
    x = torch.randn(rows, C, device=dev).half()
    g, bt = torch.randn(C, device=dev), torch.randn(C, device=dev)
    outs = [torch.empty(rows, C, device=dev, dtype=torch.float16) for _ in range(2)]
    prog = ops.Program()
    for k in range(reps):
        ops.layernorm(prog, x, outs[k & 1], g, bt)
    us = timed(prog)
    print(f"layernorm rows={rows} C={C}: {us:7.2f} us  {4.0 * rows * C / us / 1e3:7.1f} GB/s")
