#!/usr/bin/env python
"""Micro-benchmark of the fused attention kernel over the shapes of the SD-1.x dual-stream step (CUDA events,
L2-cold inputs rotated over several buffers)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from uni_renderer_b200 import ops  # noqa: E402

shapes = [(4, 8, 4096, 4096, 40), (4, 8, 1024, 1024, 80), (4, 8, 256, 256, 160), (4, 8, 4096, 77, 40),
          (4, 8, 1024, 77, 80), (2, 8, 16384, 16384, 40)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
g = torch.Generator(device="cuda").manual_seed(0)
for (B, H, Nq, Nk, d) in shapes:
    C = H * d
    nbuf = 4
    qs = [torch.randn(B * Nq, C, generator=g, device="cuda").half() for _ in range(nbuf)]
    ks = [torch.randn(B * Nk, C, generator=g, device="cuda").half() for _ in range(nbuf)]
    vs = [torch.randn(B * Nk, C, generator=g, device="cuda").half() for _ in range(nbuf)]
    out = torch.empty(B * Nq, C, device="cuda", dtype=torch.half)
    for i in range(3):
        ops.attention(None, qs[i % nbuf], ks[i % nbuf], vs[i % nbuf], out, B=B, heads=H, Nq=Nq, Nk=Nk, d=d)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ops.attention(None, qs[i % nbuf], ks[i % nbuf], vs[i % nbuf], out, B=B, heads=H, Nq=Nq, Nk=Nk, d=d)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    fl = 4.0 * B * H * Nq * Nk * d
    exps = B * H * Nq * Nk
    print(f"attn B={B} H={H} Nq={Nq} Nk={Nk} d={d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
          f"{exps / us / 1e3 / 148 / 1.965:6.2f} exp/clk/SM")
