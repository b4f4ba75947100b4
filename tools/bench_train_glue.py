#!/usr/bin/env python
"""Micro-benchmark + bit-exactness check of the training glue kernels (unib200_pack_master_weight, unib200_wgrad_scatter_add)
against the torch permute / flip / pad / cast chains they replace (ops.pack_weight, train.dgrad_weight, permute + add)."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uni_renderer_b200 import ops, train as T
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
for (O, I, k) in [(1280, 1280, 3), (320, 320, 3), (1280, 2560, 3), (1280, 1280, 1), (320, 7, 3), (4, 320, 3), (200, 77, 1), (640, 1000, 3)]:
    w = torch.randn(O, I, k, k, device="cuda") if k == 3 else torch.randn(O, I, device="cuda")
    taps = k * k
    bf = torch.zeros(O, taps * ((I + 63) // 64 * 64), device="cuda", dtype=torch.half)
    bd = torch.zeros(I, taps * ((O + 63) // 64 * 64), device="cuda", dtype=torch.half)
    t_f = timeit(lambda: T.pack_master_weight(w, O, I, taps, bf, False))
    t_d = timeit(lambda: T.pack_master_weight(w, O, I, taps, bd, True))
    kind = ops.SEG_3x3 if k == 3 else ops.SEG_1x1
    t_tf = timeit(lambda: ops.pack_weight([(w, kind)]))
    t_td = timeit(lambda: T.dgrad_weight(w))
    assert torch.equal(bf, ops.pack_weight([(w, kind)])) and torch.equal(bd, T.dgrad_weight(w))
    dw = torch.randn(O, taps, I, device="cuda"); g = torch.zeros(O, I, k, k, device="cuda") if k == 3 else torch.zeros(O, I, device="cuda")
    import ctypes as C
    from uni_renderer_b200 import _lib as L
    lib = L.load()
    t_s = timeit(lambda: L.check(lib.unib200_wgrad_scatter_add(None, dw.data_ptr(), g.data_ptr(), O, taps, I, torch.cuda.current_stream().cuda_stream), "x"))
    g.zero_(); L.check(lib.unib200_wgrad_scatter_add(None, dw.data_ptr(), g.data_ptr(), O, taps, I, torch.cuda.current_stream().cuda_stream), "x")
    assert torch.equal(g.reshape(-1), dw.reshape(O, k, k, I).permute(0, 3, 1, 2).contiguous().reshape(-1)), "scatter-add mismatch"
    t_ts = timeit(lambda: g.add_(dw.reshape(O, k, k, I).permute(0, 3, 1, 2).contiguous().reshape(g.shape)))
    print(f"O={O} I={I} k={k}: pack fwd native {t_f:.1f} us torch {t_tf:.1f} | dgrad native {t_d:.1f} torch {t_td:.1f} | scatter-add native {t_s:.1f} torch {t_ts:.1f}")
