"""Torch-tensor front end of the C ABI: builds the descriptor structs from tensors (device pointers, leading dims)
and calls libunib200.so.  PyTorch is used for device memory and streams only; no torch compute op is on this path.
Every function takes `prog` (a recorded program handle or None for an immediate launch)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import (EPI_AXPBY, EPI_GEGLU, EPI_OUT_F32, EPI_OUT_NCHW, EPI_SILU, SEG_1x1, SEG_3x3, SEG_3x3_S2,
                   SEG_3x3_S2P0, SEG_UP2x2)

__all__ = ["conv_gemm", "conv_gemm_dual", "attention", "groupnorm", "layernorm", "to_nhwc", "from_nhwc", "upsample2x",
           "timestep_sinusoid", "gemv", "axpby", "unipc_step", "add_int", "add_f16", "softmax_rows", "gaussian_sample", "Program", "Context", "pack_weight", "pack_geglu", "fold_layernorm", "rowstats_parts", "device_info",
           "SEG_1x1", "SEG_3x3", "SEG_3x3_S2", "SEG_3x3_S2P0", "SEG_UP2x2", "pack_upsample_conv", "upfold_supported", "EPI_GEGLU", "EPI_OUT_NCHW", "EPI_OUT_F32", "EPI_SILU", "EPI_AXPBY"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _check_2d(t: torch.Tensor, what: str) -> int:
    """fp16 CUDA matrix with unit column stride; returns its leading dimension (row stride in elements)."""
    if not (t.is_cuda and t.dtype == torch.float16 and t.dim() == 2 and t.stride(1) == 1):
        raise ValueError(f"{what}: expected a CUDA fp16 [rows, cols] tensor with unit column stride, got "
                         f"{tuple(t.shape)} {t.dtype} strides {t.stride()} on {t.device}")
    return t.stride(0)


_pdl_mode = [None]


def set_pdl(mode: int) -> int:
    """Programmatic dependent launch for the kernels launched / captured from now on (0 off, 1 trigger at CTA end,
    2 early trigger); returns the previous mode.  An explicit UNIB200_PDL (A/B runs) pins the mode: calls are ignored."""
    import os
    if _pdl_mode[0] is None:
        _pdl_mode[0] = int(os.environ.get("UNIB200_PDL", "0"))
    prev = _pdl_mode[0]
    if os.environ.get("UNIB200_PDL") is None and mode != prev:
        L.load().unib200_set_pdl(int(mode))
        _pdl_mode[0] = int(mode)
    return prev


def device_info() -> Tuple[int, int, int]:
    lib = L.load()
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    L.check(lib.unib200_device_info(C.byref(a), C.byref(b), C.byref(c)), "device_info")
    return a.value, b.value, c.value


class Program:
    """A recorded op list (include/unib200.h unib200_program): replayed op by op or as one CUDA graph."""

    def __init__(self):
        self.lib = L.load()
        self.handle = self.lib.unib200_program_create()
        if not self.handle:
            raise L.Unib200Error("program_create failed")
        self._keep: list = []          # tensors referenced by recorded ops
        self.has_graph = False

    def keep(self, *tensors):
        self._keep.extend(t for t in tensors if t is not None)

    @property
    def num_launches(self) -> int:
        return self.lib.unib200_program_num_launches(self.handle)

    def run(self):
        L.check(self.lib.unib200_program_run(self.handle, _stream()), "program_run")

    def lane(self, n: int):
        """Ops recorded from now on run on lane n (0 = caller's stream; lanes run concurrently between barriers)."""
        L.check(self.lib.unib200_program_set_lane(self.handle, n), "program_set_lane")

    def barrier(self):
        """Join every lane into lane 0, then fork again."""
        L.check(self.lib.unib200_program_barrier(self.handle), "program_barrier")

    def instantiate_graph(self):
        L.check(self.lib.unib200_program_graph_instantiate(self.handle, _stream()), "graph_instantiate")
        self.has_graph = True

    def launch_graph(self):
        L.check(self.lib.unib200_program_graph_launch(self.handle, _stream()), "graph_launch")

    @property
    def num_ops(self) -> int:
        return self.lib.unib200_program_num_ops(self.handle)

    def op_info(self):
        """[(kind, flops, bytes, launches)] per recorded op (algorithmic counts, see include/unib200.h)."""
        out = []
        k, f, b, n = C.c_int(), C.c_double(), C.c_double(), C.c_int()
        for i in range(self.num_ops):
            L.check(self.lib.unib200_program_op_info(self.handle, i, C.byref(k), C.byref(f), C.byref(b), C.byref(n)),
                    "op_info")
            out.append((k.value, f.value, b.value, n.value))
        return out

    def op_desc(self, i: int) -> str:
        return self.lib.unib200_program_op_desc(self.handle, i).decode()

    def profile(self, iters: int = 3):
        """Mean device ms of every op (CUDA events around each op on the current stream); host-synchronous."""
        n = self.num_ops
        buf = (C.c_float * n)()
        L.check(self.lib.unib200_program_profile(self.handle, _stream(), iters, buf), "program_profile")
        return list(buf)

    def __del__(self):
        try:
            if self.handle and getattr(self, "_owned", True):
                self.lib.unib200_program_destroy(self.handle)
            self.handle = None
        except Exception:
            pass


class Context:
    """Step-level context of the C ABI (include/unib200.h unib200_ctx): owns the programs attached to it and the buffer
    bindings of one sampling plan, so that a whole sampling loop is ONE C call (`sample_loop`) -- no Python inside it."""

    def __init__(self, device_index: int, use_graph: bool = True):
        self.lib = L.load()
        cfg = (C.c_int * 8)(int(use_graph), 0, 0, 0, 0, 0, 0, 0)
        self.handle = self.lib.unib200_create(device_index, cfg)
        if not self.handle:
            raise L.Unib200Error("unib200_create failed")
        self._keep: list = []

    def attach(self, name: str, prog: Program):
        """Ownership of the program passes to the context (the Python handle stays usable while the context lives)."""
        L.check(self.lib.unib200_ctx_attach(self.handle, name.encode(), prog.handle), "ctx_attach")
        prog._owned = False
        self._keep.append(prog)

    def bind(self, key: str, t: torch.Tensor):
        assert t.is_cuda and t.is_contiguous()
        L.check(self.lib.unib200_bind(self.handle, key.encode(), t.data_ptr(), t.numel() * t.element_size()), "ctx_bind")
        self._keep.append(t)

    def run(self, name: str):
        L.check(self.lib.unib200_ctx_run(self.handle, name.encode(), _stream()), "ctx_run")

    def dual_step(self):
        L.check(self.lib.unib200_dual_step(self.handle, _stream()), "dual_step")

    def sample_loop(self, n_steps: int, lat_img=None, lat_attr=None, ehs=None, out_img=None, out_attr=None):
        """Pointers (ints) or None: host or device memory of exactly the bound buffers' sizes."""
        L.check(self.lib.unib200_sample_loop(self.handle, n_steps, lat_img, lat_attr, ehs, out_img, out_attr, _stream()),
                "sample_loop")

    def __del__(self):
        try:
            if self.handle:
                self.lib.unib200_destroy(self.handle)
                self.handle = None
                for p in self._keep:
                    if isinstance(p, Program):
                        p.handle = None
        except Exception:
            pass


def _h(prog: Optional[Program]):
    return prog.handle if prog is not None else None


# ---------------------------------------------------------------------------------------------------------------
# weight packing (one-time, at load): reference layout (diffusers state dict) -> [N, Ktot] fp16 K-major
# ---------------------------------------------------------------------------------------------------------------
def pack_weight(parts: Sequence[Tuple[torch.Tensor, int]], device=None) -> torch.Tensor:
    """parts: list of (weight, kind) accumulated into the same output, in K order.
    weight is [O, I] (linear), [O, I, 1, 1] or [O, I, 3, 3]; kind is SEG_*.  Channels are padded to a multiple of
    64 per tap, taps are row-major (dy, dx) -- the order the kernel's K loop walks (csrc/gemm_sm100.cu)."""
    cols = []
    for w, kind in parts:
        w = w.detach().float()
        if w.dim() == 2:
            w = w[:, :, None, None]
        O, I, kh, kw = w.shape
        if kind == SEG_1x1:
            assert kh == 1 and kw == 1
        else:
            assert kh == 3 and kw == 3
        Ipad = (I + 63) // 64 * 64
        wp = torch.zeros(O, kh * kw, Ipad, dtype=torch.float32, device=w.device)
        wp[:, :, :I] = w.permute(0, 2, 3, 1).reshape(O, kh * kw, I)
        cols.append(wp.reshape(O, kh * kw * Ipad))
    out = torch.cat(cols, dim=1).to(torch.float16).contiguous()
    return out.to(device) if device is not None else out


_UP_TAPS = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}     # (parity, 2x2 tap) -> 3x3 taps it sums


def pack_upsample_conv(w: torch.Tensor) -> torch.Tensor:
    """Upsample2D = nearest-2x + conv3x3 [O, I, 3, 3] as four 2x2 convs on the LOW-resolution input (SEG_UP2x2,
    include/unib200.h): output pixel (2h+py, 2w+px) sees input pixels (h-1+py+ty, w-1+px+tx), ty, tx in {0, 1}; its
    weights are the sums of the 3x3 taps that land on the same input pixel.  Returns fp16 [4 * O, 4 * ceil(I/64)*64],
    rows ordered [parity py*2+px][O], K ordered [tap ty*2+tx][channel].  The sums are formed in fp32 and rounded ONCE."""
    w = w.detach().float()
    O, I, kh, kw = w.shape
    assert kh == 3 and kw == 3
    Ipad = (I + 63) // 64 * 64
    out = torch.zeros(4, O, 4, Ipad, dtype=torch.float32, device=w.device)
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    acc = 0
                    for dy in _UP_TAPS[(py, ty)]:
                        for dx in _UP_TAPS[(px, tx)]:
                            acc = acc + w[:, :, dy, dx]
                    out[py * 2 + px, :, ty * 2 + tx, :I] = acc
    return out.reshape(4 * O, 4 * Ipad).to(torch.float16).contiguous()


def upfold_supported(cout: int) -> bool:
    """SEG_UP2x2 needs Cout to be a power-of-two multiple of the N tile the kernel picks for N = 4 * Cout."""
    bn = pick_bn(4 * cout)
    t = cout // bn if bn and cout % bn == 0 else 0
    return t >= 1 and (t & (t - 1)) == 0 and cout % 32 == 0


def pick_bn(N: int, flags: int = 0) -> int:
    """N-tile width the kernel uses for this (N, flags) -- asked from the library so packing always agrees."""
    return L.load().unib200_pick_bn(N, flags)


def rowstats_parts(N: int, flags: int = 0) -> int:
    """Row-statistics partials a GEMM with N output channels writes per row: two per N tile (the two epilogue warps
    that share a row of the tile each sum their own column sub-tiles, csrc/gemm_sm100.cu)."""
    bn = pick_bn(N, flags)
    return 2 * ((N + bn - 1) // bn)


def fold_layernorm(w: torch.Tensor, bias: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm(x) @ w.T + bias  ==  rstd * (x @ (w * gamma).T - mean * wsum) + bias2  (see include/unib200.h).
    Returns (w * gamma as fp16-representable fp32, wsum fp32 [N], bias2 fp32 [N]); wsum is taken over the
    fp16-ROUNDED folded weights, i.e. exactly what the tensor core sums."""
    w32, g32, b32 = w.detach().float(), gamma.detach().float(), beta.detach().float()
    wf = (w32 * g32[None, :]).half().float()
    wsum = wf.sum(dim=1).contiguous()
    bias2 = (w32 @ b32)
    if bias is not None:
        bias2 = bias2 + bias.detach().float()
    return wf, wsum, bias2.contiguous()


def pack_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU projection [2*inner, C] (+bias): interleave value/gate rows per N-tile so one accumulator tile holds
    both halves of the same output columns (EPI_GEGLU)."""
    two_inner = w.shape[0]
    inner = two_inner // 2
    bn = pick_bn(two_inner, EPI_GEGLU)
    half = bn // 2
    assert bn > 0 and two_inner % bn == 0 and half % 32 == 0 and inner % half == 0
    idx = []
    for t in range(two_inner // bn):
        idx.append(torch.arange(t * half, (t + 1) * half))
        idx.append(inner + torch.arange(t * half, (t + 1) * half))
    idx = torch.cat(idx).to(w.device)
    return w[idx].contiguous(), b[idx].contiguous()


# ---------------------------------------------------------------------------------------------------------------
# ops
# ---------------------------------------------------------------------------------------------------------------
def conv_gemm(prog: Optional[Program], segs: Sequence[Tuple[torch.Tensor, int, int]], weight: torch.Tensor,
              out: torch.Tensor, *, M: int, N: int, B: int = 0, H: int = 0, W: int = 0,
              bias: Optional[torch.Tensor] = None, bias_bstride: int = 0, bias_step: Optional[torch.Tensor] = None,
              bias_step_stride: int = 0, res: Optional[torch.Tensor] = None,
              flags: int = 0, splits: int = 0, partial: Optional[torch.Tensor] = None,
              axpby: Optional[torch.Tensor] = None, axpby_step: Optional[torch.Tensor] = None,
              aux: Optional[torch.Tensor] = None, aux_out: Optional[torch.Tensor] = None,
              axpby_first_channel: int = 0, ldc: Optional[int] = None, rowstats_out: Optional[torch.Tensor] = None,
              ln: Optional[Tuple[torch.Tensor, torch.Tensor, float, int]] = None,
              gn: Optional[Tuple[torch.Tensor, int, int]] = None):
    """segs: (matrix [pixels, >=C] fp16, C, kind).  H = W = 0 selects the plain row-major [M, K] path.
    gn = (part fp32 [M / rows, N / gran, 2], gran, rows): also emit the GroupNorm statistics of the output."""
    d, keep = _gemm_desc(segs, weight, out, M=M, N=N, B=B, H=H, W=W, bias=bias, bias_bstride=bias_bstride,
                         bias_step=bias_step, bias_step_stride=bias_step_stride, res=res, flags=flags, splits=splits,
                         partial=partial, axpby=axpby, axpby_step=axpby_step, aux=aux, aux_out=aux_out,
                         axpby_first_channel=axpby_first_channel, ldc=ldc, rowstats_out=rowstats_out, ln=ln, gn=gn)
    L.check(L.load().unib200_conv_gemm(_h(prog), C.byref(d), _stream()), "conv_gemm")
    if prog is not None:
        prog.keep(*keep)


def conv_gemm_dual(prog: Optional[Program], a: dict, b: dict):
    """Two conv_gemm problems of identical shape (dicts of conv_gemm's arguments: segs, weight, out, M, N, B, bias, res,
    gn) as ONE kernel launch (include/unib200.h unib200_conv_gemm_dual) -- both directions of the dual-stream exchange."""
    da, ka = _gemm_desc(**a)
    db, kb = _gemm_desc(**b)
    L.check(L.load().unib200_conv_gemm_dual(_h(prog), C.byref(da), C.byref(db), _stream()), "conv_gemm_dual")
    if prog is not None:
        prog.keep(*ka, *kb)


def _gemm_desc(segs, weight, out, *, M, N, B=0, H=0, W=0, bias=None, bias_bstride=0, bias_step=None,
               bias_step_stride=0, res=None, flags=0, splits=0, partial=None, axpby=None, axpby_step=None, aux=None,
               aux_out=None, axpby_first_channel=0, ldc=None, rowstats_out=None, ln=None, gn=None):
    """unib200_gemm_desc from tensors + the list of tensors a recorded program must keep alive."""
    d = L.GemmDesc()
    d.M, d.N, d.B, d.H, d.W, d.nseg = M, N, B, H, W, len(segs)
    ktot = 0
    for i, (t, c, kind) in enumerate(segs):
        ld = _check_2d(t, f"conv_gemm seg {i}")
        d.seg[i].ptr, d.seg[i].C, d.seg[i].ld, d.seg[i].kind = t.data_ptr(), c, ld, kind
        ktot += (1 if kind == SEG_1x1 else 4 if kind == SEG_UP2x2 else 9) * ((c + 63) // 64 * 64)
    if weight.dtype != torch.float16 or tuple(weight.shape) != (N, ktot) or not weight.is_contiguous():
        raise ValueError(f"conv_gemm: packed weight must be contiguous fp16 [{N}, {ktot}], got {tuple(weight.shape)}")
    d.weight = weight.data_ptr()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.stride(-1) == 1
    d.bias, d.bias_bstride = _ptr(bias), bias_bstride
    if bias_step is not None:
        assert bias_step.dtype == torch.int32
        d.bias_step, d.bias_step_stride = bias_step.data_ptr(), bias_step_stride
    if res is not None:
        d.ldr = _check_2d(res, "conv_gemm res")
    d.res = _ptr(res)
    d.out = _ptr(out)
    if ldc is not None:
        d.ldc = ldc
    elif out is not None and not (flags & EPI_OUT_NCHW):
        d.ldc = _check_2d(out, "conv_gemm out")
    d.flags, d.splits = flags, splits
    if partial is not None:
        d.partial, d.partial_bytes = partial.data_ptr(), partial.numel() * partial.element_size()
    d.axpby, d.axpby_step, d.aux, d.aux_out = _ptr(axpby), _ptr(axpby_step), _ptr(aux), _ptr(aux_out)
    d.axpby_first_channel = axpby_first_channel
    if rowstats_out is not None:      # producer of a LayerNorm input: [M, 2 * ceil(N / BN), 2] fp32 row statistics
        assert rowstats_out.dtype == torch.float32 and rowstats_out.is_contiguous()
        assert rowstats_out.numel() >= M * rowstats_parts(N, flags) * 2
        d.rowstats_out = rowstats_out.data_ptr()
    ln_keep = ()
    if ln is not None:                # consumer of LayerNorm(x): (rowstats [M, parts, 2], wsum [N], eps, C)
        rs, wsum, eps, Cn = ln
        assert rs.dtype == torch.float32 and wsum.dtype == torch.float32 and wsum.numel() == N and rs.shape[0] == M
        d.ln_rowstats, d.ln_parts, d.ln_wsum, d.ln_eps, d.ln_C = rs.data_ptr(), rs.shape[1], wsum.data_ptr(), eps, Cn
        ln_keep = (rs, wsum)
    if gn is not None:
        part, gran, rows = gn
        assert part.dtype == torch.float32 and part.is_contiguous() and part.numel() >= (M // rows) * (N // gran) * 2
        d.gn_part, d.gn_gran, d.gn_rows = part.data_ptr(), gran, rows
    return d, (*(s[0] for s in segs), weight, out, bias, bias_step, res, partial, axpby, axpby_step, aux, aux_out,
               rowstats_out, *ln_keep, gn[0] if gn is not None else None)


def attention(prog: Optional[Program], q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *,
              B: int, heads: int, Nq: int, Nk: int, d: int, scale: Optional[float] = None,
              lse2: Optional[torch.Tensor] = None):
    """lse2 (optional, fp32 [B, heads, Nq]): receives every row's log2-domain log-sum-exp -- the training path's
    flash backward (train.attention_backward) needs it."""
    lib = L.load()
    a = L.AttnDesc()
    if lse2 is not None:
        assert lse2.dtype == torch.float32 and lse2.is_contiguous() and lse2.numel() == B * heads * Nq
        a.lse2 = lse2.data_ptr()
    a.q, a.ldq = q.data_ptr(), _check_2d(q, "attention q")
    a.k, a.ldk = k.data_ptr(), _check_2d(k, "attention k")
    a.v, a.ldv = v.data_ptr(), _check_2d(v, "attention v")
    a.out, a.ldo = out.data_ptr(), _check_2d(out, "attention out")
    a.B, a.heads, a.Nq, a.Nk, a.d = B, heads, Nq, Nk, d
    a.scale = float(scale if scale is not None else d ** -0.5)
    L.check(lib.unib200_attention(_h(prog), C.byref(a), _stream()), "attention")
    if prog is not None:
        prog.keep(q, k, v, out, lse2)


def groupnorm(prog: Optional[Program], x1: torch.Tensor, C1: int, x2: Optional[torch.Tensor], C2: int,
              gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, scratch: torch.Tensor, *, B: int, HW: int,
              groups: int, eps: float, silu: bool,
              parts: Optional[Tuple[torch.Tensor, Optional[torch.Tensor], int, int]] = None):
    """parts = (part1, part2 | None, gran, rows): statistics already written by the epilogues of the GEMMs that produced
    x1 / x2 (conv_gemm(..., gn=...)) -- then this is ONE apply launch."""
    lib = L.load()
    g = L.GnDesc()
    if parts is not None:
        p1, p2, gran, rows = parts
        assert p1.dtype == torch.float32 and (x2 is None or p2 is not None)
        g.part1, g.part2, g.part_gran, g.part_rows = p1.data_ptr(), _ptr(p2), gran, rows
    g.x1, g.ld1, g.C1 = x1.data_ptr(), _check_2d(x1, "groupnorm x1"), C1
    if x2 is not None:
        g.x2, g.ld2, g.C2 = x2.data_ptr(), _check_2d(x2, "groupnorm x2"), C2
    g.B, g.HW, g.groups, g.eps = B, HW, groups, eps
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and scratch.dtype == torch.float32
    g.gamma, g.beta, g.out, g.silu = gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), int(silu)
    assert out.is_contiguous() and out.shape[-1] == C1 + (C2 if x2 is not None else 0)
    g.scratch, g.scratch_floats = scratch.data_ptr(), scratch.numel()
    L.check(lib.unib200_groupnorm(_h(prog), C.byref(g), _stream()), "groupnorm")
    if prog is not None:
        prog.keep(x1, x2, gamma, beta, out, scratch, *(parts[:2] if parts is not None else ()))


def layernorm(prog: Optional[Program], x: torch.Tensor, y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
              eps: float = 1e-5):
    lib = L.load()
    assert x.is_contiguous() and y.is_contiguous() and x.dtype == torch.float16 and y.dtype == torch.float16
    rows, Cn = x.shape
    L.check(lib.unib200_layernorm(_h(prog), x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rows, Cn,
                                  eps, _stream()), "layernorm")
    if prog is not None:
        prog.keep(x, y, gamma, beta)


def to_nhwc(prog: Optional[Program], src: torch.Tensor, dst: torch.Tensor, Cpad: int):
    """src: logical [B, C, H, W] fp32/fp16 with arbitrary strides -> dst [B*H*W, Cpad] fp16 (zero padded)."""
    lib = L.load()
    assert src.dtype in (torch.float32, torch.float16) and dst.dtype == torch.float16 and dst.is_contiguous()
    B, Cn, H, W = src.shape
    sb, sc, sh, sw = src.stride()
    L.check(lib.unib200_to_nhwc(_h(prog), src.data_ptr(), int(src.dtype == torch.float32), dst.data_ptr(), B, Cn, H, W,
                                sb, sc, sh, sw, Cpad, _stream()), "to_nhwc")
    if prog is not None:
        prog.keep(src, dst)


def from_nhwc(prog: Optional[Program], src: torch.Tensor, dst: torch.Tensor, *, B: int, Cn: int, HW: int):
    """src [B*HW, ld] fp16 (first Cn channels) -> contiguous NCHW dst (fp32 or fp16)."""
    lib = L.load()
    ld = _check_2d(src, "from_nhwc src")
    assert dst.is_contiguous() and dst.dtype in (torch.float32, torch.float16)
    L.check(lib.unib200_from_nhwc(_h(prog), src.data_ptr(), dst.data_ptr(), int(dst.dtype == torch.float32), B, Cn, HW,
                                  ld, _stream()), "from_nhwc")
    if prog is not None:
        prog.keep(src, dst)


def upsample2x(prog: Optional[Program], src: torch.Tensor, dst: torch.Tensor, *, B: int, H: int, W: int, Cn: int):
    lib = L.load()
    assert src.is_contiguous() and dst.is_contiguous()
    L.check(lib.unib200_upsample2x(_h(prog), src.data_ptr(), dst.data_ptr(), B, H, W, Cn, _stream()), "upsample2x")
    if prog is not None:
        prog.keep(src, dst)


def timestep_sinusoid(prog: Optional[Program], t: torch.Tensor, out: torch.Tensor, *, B: int, dim: int,
                      step_idx: Optional[torch.Tensor] = None, t_stride: int = 0):
    lib = L.load()
    assert t.dtype == torch.float32 and out.dtype == torch.float32
    L.check(lib.unib200_timestep_sinusoid(_h(prog), t.data_ptr(), _ptr(step_idx), t_stride, out.data_ptr(), B, dim,
                                          _stream()), "timestep_sinusoid")
    if prog is not None:
        prog.keep(t, out, step_idx)


def gemv(prog: Optional[Program], x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], y: torch.Tensor, *,
         silu: bool):
    lib = L.load()
    B, K = x.shape
    N = w.shape[0]
    assert x.dtype == torch.float32 and w.dtype == torch.float16 and y.dtype == torch.float32
    assert x.is_contiguous() and w.is_contiguous() and y.is_contiguous() and w.shape[1] == K and y.shape == (B, N)
    L.check(lib.unib200_gemv(_h(prog), x.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), B, K, N, int(silu),
                             _stream()), "gemv")
    if prog is not None:
        prog.keep(x, w, bias, y)


def axpby(prog: Optional[Program], model_out: torch.Tensor, x: torch.Tensor, out: torch.Tensor, coef: torch.Tensor,
          step_idx: Optional[torch.Tensor] = None):
    lib = L.load()
    assert model_out.dtype == x.dtype == out.dtype == coef.dtype == torch.float32
    assert model_out.is_contiguous() and x.is_contiguous() and out.is_contiguous()
    L.check(lib.unib200_axpby(_h(prog), model_out.data_ptr(), x.data_ptr(), out.data_ptr(), coef.data_ptr(),
                              _ptr(step_idx), x.numel(), _stream()), "axpby")
    if prog is not None:
        prog.keep(model_out, x, out, coef, step_idx)


def unipc_step(prog: Optional[Program], model_out: torch.Tensor, sample: torch.Tensor, last_sample: torch.Tensor,
               hist0: torch.Tensor, hist1: torch.Tensor, coef: torch.Tensor, step_idx: Optional[torch.Tensor] = None,
               first_channel: int = 0):
    """One fused UniPC update of a [B, C, H, W] fp32 latent (include/unib200.h unib200_unipc_step)."""
    lib = L.load()
    B, Cn = sample.shape[0], sample.shape[1]
    HW = sample.numel() // (B * Cn)
    for t in (model_out, sample, last_sample, hist0, hist1):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.shape == sample.shape
    assert coef.dtype == torch.float32 and coef.is_contiguous() and coef.shape[-1] == 10
    L.check(lib.unib200_unipc_step(_h(prog), model_out.data_ptr(), sample.data_ptr(), last_sample.data_ptr(),
                                   hist0.data_ptr(), hist1.data_ptr(), coef.data_ptr(), _ptr(step_idx), B, Cn, HW,
                                   first_channel, _stream()), "unipc_step")
    if prog is not None:
        prog.keep(model_out, sample, last_sample, hist0, hist1, coef, step_idx)


def softmax_rows(prog: Optional[Program], s: torch.Tensor, *, rows: int, n: int, scale: float):
    """In-place P = softmax(scale * S) over the first n columns of each of `rows` rows of an fp16 matrix."""
    lib = L.load()
    ld = _check_2d(s, "softmax_rows s")
    if s.shape[0] < rows or s.shape[1] < n:
        raise ValueError(f"softmax_rows: matrix {tuple(s.shape)} is smaller than rows={rows}, n={n}")
    L.check(lib.unib200_softmax_rows(_h(prog), s.data_ptr(), rows, n, ld, float(scale), _stream()), "softmax_rows")
    if prog is not None:
        prog.keep(s)


def gaussian_sample(prog: Optional[Program], moments: torch.Tensor, noise: Optional[torch.Tensor], out: torch.Tensor, *,
                    scale: float = 1.0):
    """out = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scale from moments [B, 2C, H, W] fp32; noise None
    gives the mode (include/unib200.h unib200_gaussian_sample)."""
    lib = L.load()
    B, C2, H, W = moments.shape
    assert moments.dtype == torch.float32 and moments.is_contiguous() and moments.is_cuda and C2 % 2 == 0
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, C2 // 2, H, W)
    if noise is not None:
        assert noise.dtype == torch.float32 and noise.is_contiguous() and noise.shape == out.shape and noise.is_cuda
    L.check(lib.unib200_gaussian_sample(_h(prog), moments.data_ptr(), _ptr(noise), out.data_ptr(), B, C2 // 2, H * W,
                                        float(scale), _stream()), "gaussian_sample")
    if prog is not None:
        prog.keep(moments, noise, out)


def add_int(prog: Optional[Program], p: torch.Tensor, v: int):
    lib = L.load()
    assert p.dtype == torch.int32
    L.check(lib.unib200_add_int(_h(prog), p.data_ptr(), v, _stream()), "add_int")
    if prog is not None:
        prog.keep(p)


def add_f16(prog: Optional[Program], a: torch.Tensor, b: torch.Tensor, out: torch.Tensor):
    lib = L.load()
    assert a.dtype == b.dtype == out.dtype == torch.float16
    assert a.is_contiguous() and b.is_contiguous() and out.is_contiguous() and a.numel() == b.numel() == out.numel()
    L.check(lib.unib200_add_f16(_h(prog), a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "add_f16")
    if prog is not None:
        prog.keep(a, b, out)
