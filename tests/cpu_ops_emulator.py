"""TEST INFRASTRUCTURE: a torch-CPU emulation of the C-ABI ops' documented semantics (include/unib200.h), used to
check the HOST-SIDE wiring of recorded programs (segment order, packed-weight layout, fused shortcuts, the
GEMM-softmax-GEMM attention, the quant_conv fold) against the oracle without a GPU.

It is installed by monkeypatching `uni_renderer_b200.ops` inside a test; the product never imports it, and it is not a
fallback: nothing in uni_renderer_b200/ can reach it.  Kernel-level behaviour is covered by the `-m gpu` tests."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from uni_renderer_b200 import _lib as L


class FakeProgram:
    def __init__(self):
        self.ops = []
        self.has_graph = False

    def keep(self, *a):
        pass

    def run(self):
        for op in self.ops:
            op()

    @property
    def num_launches(self):
        return len(self.ops)


def _submit(prog, fn):
    if prog is None:
        fn()
    else:
        prog.ops.append(fn)


def _taps(x, kind, B, H, W, C):
    """x: [B*Hin*Win, >=C] -> list of [B*H*W, C] fp32 tap matrices in the kernel's K order."""
    if kind == L.SEG_1x1:
        return [x[:, :C].float()]
    s = 1 if kind == L.SEG_3x3 else 2
    Hin, Win = H * s, W * s
    img = x[:, :C].float().reshape(B, Hin, Win, C)
    if kind == L.SEG_3x3_S2P0:
        xp = F.pad(img, (0, 0, 0, 2, 0, 2))            # bottom / right only (one pixel needed, two keeps slices easy)
    else:
        xp = F.pad(img, (0, 0, 1, 1, 1, 1))
    out = []
    for dy in range(3):
        for dx in range(3):
            out.append(xp[:, dy:dy + s * H:s, dx:dx + s * W:s].reshape(B * H * W, C))
    return out


def conv_gemm(prog, segs, weight, out, *, M, N, B=0, H=0, W=0, bias=None, bias_bstride=0, bias_step=None,
              bias_step_stride=0, res=None, flags=0, splits=0, partial=None, axpby=None, axpby_step=None, aux=None,
              aux_out=None, axpby_first_channel=0, ldc=None, rowstats_out=None, ln=None):
    if flags & ~(L.EPI_OUT_NCHW | L.EPI_OUT_F32 | L.EPI_SILU) or bias_step is not None or ln is not None \
            or rowstats_out is not None:
        raise NotImplementedError("emulator: epilogue not modelled")
    ktot = sum((1 if k == L.SEG_1x1 else 9) * ((c + 63) // 64 * 64) for _, c, k in segs)
    assert weight.dtype == torch.float16 and tuple(weight.shape) == (N, ktot) and weight.is_contiguous()
    linear = (H == 0 or W == 0)

    def run():
        cols = []
        for t, c, kind in segs:
            assert t.dtype == torch.float16 and t.stride(1) == 1
            if linear:
                assert kind == L.SEG_1x1 and t.shape[0] >= M
                taps = [t[:M, :c].float()]
            else:
                taps = _taps(t, kind, B, H, W, c)
            cpad = (c + 63) // 64 * 64
            for tp in taps:
                cols.append(F.pad(tp, (0, cpad - c)))
        acc = torch.cat(cols, 1) @ weight.float().t()
        rpb = (M // B) if B > 0 else M
        if bias is not None:
            if bias_bstride:
                acc = acc + bias.reshape(-1, bias_bstride)[:, :N].float().repeat_interleave(rpb, 0)[:M]
            else:
                acc = acc + bias.float()[None, :N]
        if res is not None:
            acc = acc + res[:M, :N].float()
        if flags & L.EPI_SILU:
            acc = F.silu(acc)
        if flags & L.EPI_OUT_NCHW:
            nb = M // rpb
            out.reshape(nb, N, rpb).copy_(acc.reshape(nb, rpb, N).permute(0, 2, 1))
        else:
            out[:M, :N].copy_(acc)
    _submit(prog, run)


def groupnorm(prog, x1, C1, x2, C2, gamma, beta, out, scratch, *, B, HW, groups, eps, silu):
    def run():
        x = x1[:, :C1].float()
        if x2 is not None:
            x = torch.cat([x, x2[:, :C2].float()], 1)
        C = x.shape[1]
        y = F.group_norm(x.reshape(B, HW, C).permute(0, 2, 1), groups, gamma, beta, eps)
        if silu:
            y = F.silu(y)
        out.copy_(y.permute(0, 2, 1).reshape(B * HW, C))
    _submit(prog, run)


def upsample2x(prog, src, dst, *, B, H, W, Cn):
    def run():
        x = src.reshape(B, H, W, Cn)
        dst.copy_(x.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(B * 4 * H * W, Cn))
    _submit(prog, run)


def to_nhwc(prog, src, dst, Cpad):
    def run():
        B, C, H, W = src.shape
        dst.zero_()
        dst[:, :C].copy_(src.permute(0, 2, 3, 1).reshape(B * H * W, C))
    _submit(prog, run)


def softmax_rows(prog, s, *, rows, n, scale):
    def run():
        s[:rows, :n].copy_(torch.softmax(s[:rows, :n].float() * scale, dim=-1))
    _submit(prog, run)


def gaussian_sample(prog, moments, noise, out, *, scale=1.0):
    def run():
        mean, logvar = moments.chunk(2, 1)
        y = mean if noise is None else mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise
        out.copy_(y * scale)
    _submit(prog, run)


def add_f16(prog, a, b, out):
    _submit(prog, lambda: out.copy_(a.float() + b.float()))


def install(monkeypatch):
    """Route uni_renderer_b200.ops through the emulator for the duration of one test."""
    from uni_renderer_b200 import ops
    for name in ("conv_gemm", "groupnorm", "upsample2x", "to_nhwc", "softmax_rows", "gaussian_sample", "add_f16"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, "Program", FakeProgram)
