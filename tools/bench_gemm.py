#!/usr/bin/env python
"""Micro-benchmark of representative conv/linear shapes of the denoising step through the C ABI (unib200_conv_gemm).
Each shape is recorded `--reps` times into one program, replayed as a CUDA graph and timed with CUDA events, so the
figure is the steady-state per-launch time in a dependent chain (what the step graph sees).  `--only i --eager` runs one
shape a few times without a graph (for `ncu --set full -k regex:gemm_tcgen05`).  Optimisation instrument, not a
bench value."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uni_renderer_b200 import ops  # noqa: E402
from uni_renderer_b200.ops import EPI_GEGLU, SEG_1x1, SEG_3x3, SEG_3x3_S2  # noqa: E402

# (name, B, H(out), Cin list, kind, Cout, res, geglu)
SHAPES = [
    ("proj 1x1 320 @64 +res", 4, 64, [320], SEG_1x1, 320, True, False),
    ("proj 1x1 640 @32 +res", 4, 32, [640], SEG_1x1, 640, True, False),
    ("proj 1x1 1280 @16 +res", 4, 16, [1280], SEG_1x1, 1280, True, False),
    ("proj 1x1 1280 @8 +res", 4, 8, [1280], SEG_1x1, 1280, True, False),
    ("qkv 320->960 @64", 4, 64, [320], SEG_1x1, 960, False, False),
    ("geglu 320->2560 @64", 4, 64, [320], SEG_1x1, 2560, False, True),
    ("geglu 640->5120 @32", 4, 32, [640], SEG_1x1, 5120, False, True),
    ("geglu 1280->10240 @16", 4, 16, [1280], SEG_1x1, 10240, False, True),
    ("qkv 640->1920 @32", 4, 32, [640], SEG_1x1, 1920, False, False),
    ("ff2 1280->320 @64 +res", 4, 64, [1280], SEG_1x1, 320, True, False),
    ("ff2 2560->640 @32 +res", 4, 32, [2560], SEG_1x1, 640, True, False),
    ("ff2 5120->1280 @16 +res", 4, 16, [5120], SEG_1x1, 1280, True, False),
    ("conv3 320->320 @64", 4, 64, [320], SEG_3x3, 320, False, False),
    ("conv3 640->640 @32", 4, 32, [640], SEG_3x3, 640, False, False),
    ("conv3 1280->1280 @16", 4, 16, [1280], SEG_3x3, 1280, False, False),
    ("conv3 1280->1280 @8", 4, 8, [1280], SEG_3x3, 1280, False, False),
    ("conv3 2560->1280 @8", 4, 8, [2560], SEG_3x3, 1280, False, False),
    ("conv3 s2 320->320 @32", 4, 32, [320], SEG_3x3_S2, 320, False, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--eager", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    partial = torch.empty(16 << 20, device=dev, dtype=torch.float32)
    print(f"{'shape':28s} {'us':>8s} {'TF/s':>8s} {'GB/s':>8s}  desc")
    for i, (name, B, H, cins, kind, cout, res, geglu) in enumerate(SHAPES):
        if a.only >= 0 and i != a.only:
            continue
        M = B * H * H
        Hin = 2 * H if kind == SEG_3x3_S2 else H
        srcs = [torch.randn(B * Hin * Hin, c, device=dev).half() for c in cins]
        w = ops.pack_weight([(torch.randn(cout, c, *((1, 1) if kind == SEG_1x1 else (3, 3)), device=dev) * 0.02, kind)
                             for c in cins])
        bias = torch.randn(cout, device=dev)
        flags = 0
        n_out = cout
        if geglu:
            flags, n_out = EPI_GEGLU, cout // 2
        # ping-pong outputs so consecutive launches form a chain like the real step
        outs = [torch.empty(M, n_out, device=dev, dtype=torch.float16) for _ in range(2)]
        r = torch.randn(M, n_out, device=dev).half() if res else None
        prog = ops.Program()
        for k in range(a.reps):
            ops.conv_gemm(prog, [(s, c, kind) for s, c in zip(srcs, cins)], w, outs[k & 1], M=M, N=cout, B=B,
                          H=0 if kind == SEG_1x1 else H, W=0 if kind == SEG_1x1 else H, bias=bias, res=r, flags=flags,
                          partial=None if geglu else partial)
        prog.run()
        torch.cuda.synchronize()
        if a.eager:
            for _ in range(2):
                prog.run()
            torch.cuda.synchronize()
            continue
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            prog.instantiate_graph()
            side.synchronize()
            for _ in range(3):
                prog.launch_graph()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                prog.launch_graph()
            e1.record()
            side.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (a.iters * a.reps)
        kind_, fl, by, nl = prog.op_info()[0]
        print(f"{name:28s} {us:8.2f} {fl / us / 1e6:8.1f} {by / us / 1e3:8.1f}  {prog.op_desc(0)}")


if __name__ == "__main__":
    main()
