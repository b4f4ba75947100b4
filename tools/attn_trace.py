#!/usr/bin/env python
"""Debug: clock64 timeline of one attention CTA (softmax warpgroups + MMA warp)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uni_renderer_b200 import ops, _lib
B, H, N, d = 4, 8, 4096, 40
if len(sys.argv) > 2: N, d = int(sys.argv[1]), int(sys.argv[2])
C = H * d
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(B * N, C, generator=g, device="cuda").half()
k = torch.randn(B * N, C, generator=g, device="cuda").half()
v = torch.randn(B * N, C, generator=g, device="cuda").half()
out = torch.empty_like(q)
ops.attention(None, q, k, v, out, B=B, heads=H, Nq=N, Nk=N, d=d)
torch.cuda.synchronize()
tr = torch.zeros(4 * 16 * 8, dtype=torch.int64, device="cuda")
_lib.load().unib200_debug_set_trace(tr.data_ptr())
ops.attention(None, q, k, v, out, B=B, heads=H, Nq=N, Nk=N, d=d)
torch.cuda.synchronize()
_lib.load().unib200_debug_set_trace(None)
t = tr.cpu().reshape(4, 16, 8)
t0 = int(t[0, 0, 0])
names = ["WG0", "WG1", "MMA(t0)", "MMA(t1)"]
for who in range(4):
    print(names[who])
    for j in range(8):
        row = [int(x) - t0 if x else 0 for x in t[who, j, :6]]
        print("  j=%d" % j, row)
