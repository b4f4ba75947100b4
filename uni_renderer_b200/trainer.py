"""Network-level training step: forward AND backward of the reference's 3-call dual-stream step
(train/train.py:1324-1427 -- AttributeEncoderModel -> UNet2DConditionModel with the encoder's residuals ->
AttributeDecoderModel with the UNet's raw features; MSE + contrastive losses; `accelerator.backward`; clip; AdamW) on
the B200 kernels, plus the bucketed gradient all-reduce of data-parallel training.

Structure: a reverse-mode TAPE over the kernel-level operations of train.py.  Every activation is an NHWC fp16 matrix
[B*H*W, C] (`TT`); every op launches libunib200.so kernels for its forward, records a closure for its backward, and the
three networks are written against those ops exactly the way oracle/uni_oracle.py restates the reference
(models/controlnet.py:781-1166, 1657-1778, 2342-2527; models/unet_2d_blocks.py).  What runs where:

  conv3x3 / conv1x1 / linear      forward unib200_conv_gemm; dX = the same kernel on flipped / transposed weights;
                                  dW = unib200_conv_wgrad (tcgen05, pixels as the contraction dimension); db = column sums
  Downsample2D (3x3 stride 2)     forward SEG_3x3_S2; gradients through unib200_scatter2x (zero insertion) + the stride-1 kernels
  Upsample2D (nearest 2x + conv)  unib200_upsample2x / unib200_pool2x2_sum around the conv
  GroupNorm(+SiLU), LayerNorm     unib200_groupnorm(_backward), unib200_layernorm(_backward)
  attention                       forward: the flash kernel of the inference path (unib200_attention, which also emits the row
                                  log-sum-exps); backward: unib200_attention_backward, a tcgen05 flash backward (head dims
                                  <= 80: S / dP recomputed per tile, P and dS never leave the SM); wider heads (the 8x8 / 16x16
                                  levels) recompute per (sample, head) in materialised form on the GEMM / wgrad kernels
  GEGLU, SiLU, adds               unib200_geglu, unib200_silu_f16, unib200_add_f16
  time embedding MLP              the same linear ops on a 128-row padded matrix
  optimizer                       unib200_adamw_step on ONE flat fp32 parameter / gradient / moment buffer per trainer
  gradient all-reduce             torch.distributed.all_reduce over fixed-size buckets of the flat gradient buffer (NCCL on
                                  GPUs; averaged) -- `allreduce_gradients`
The loss heads (two MSE means and the three-way cosine contrastive term over a few thousand prediction values,
train/train.py:1356-1372) are evaluated by torch autograd on the fp32 predictions: that is host-level glue, not a kernel
of the path.  Activations and activation gradients are fp16 (the reference trains under fp16 autocast with a GradScaler):
`loss_scale` multiplies the seed gradients and is divided out inside the optimizer kernel.

No CPU fallback: constructing a trainer without CUDA raises."""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from . import train as T
from .ops import SEG_1x1, SEG_3x3, SEG_3x3_S2

SD = Dict[str, torch.Tensor]
_REQUIRE_CUDA = True        # tests/cpu_ops_emulator.py turns this off to drive the tape through its CPU emulation of the ops


class TT:
    """Tape tensor: fp16 matrix [B*H*W (or rows), C] + its spatial meaning + the gradient accumulated so far."""
    __slots__ = ("v", "B", "H", "W", "g", "needs_grad")

    def __init__(self, v: torch.Tensor, B: int = 1, H: int = 0, W: int = 0, needs_grad: bool = True):
        assert v.dtype == torch.float16 and v.dim() == 2 and v.is_contiguous()
        self.v, self.B, self.H, self.W, self.g, self.needs_grad = v, B, H, W, None, needs_grad

    @property
    def C(self) -> int:
        return self.v.shape[1]

    @property
    def rows(self) -> int:
        return self.v.shape[0]


class ParamSet:
    """All trainable tensors of the step in ONE flat fp32 buffer (+ one flat gradient buffer): `p[name]` / `g[name]` are
    views.  Packed fp16 kernel weights are cached per optimizer step."""

    def __init__(self, nets: Dict[str, SD], device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda" and _REQUIRE_CUDA:
            raise RuntimeError("training runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.dev = dev
        items = [(f"{net}.{k}", v) for net, sd in nets.items() for k, v in sd.items()]
        total = sum(v.numel() for _, v in items)
        self.flat = torch.empty(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.p: Dict[str, torch.Tensor] = {}
        self.g: Dict[str, torch.Tensor] = {}
        off = 0
        for name, v in items:
            n = v.numel()
            self.p[name] = self.flat[off:off + n].view(v.shape)
            self.p[name].copy_(v.detach().to(dev, torch.float32))
            self.g[name] = self.grad[off:off + n].view(v.shape)
            off += n
        self._cache: Dict[Tuple[str, str], torch.Tensor] = {}
        # persistent fp16 operand buffers of the native packing path (padding columns zeroed once, valid columns
        # rewritten by unib200_pack_master_weight after every optimizer step)
        self._packed: Dict[Tuple[str, str], torch.Tensor] = {}
        self.native_pack = dev.type == "cuda"
        self.m = self.v = None            # AdamW moments (allocated by the first step)
        self.steps = 0

    def invalidate(self):
        self._cache.clear()

    def _pack_native(self, name: str, dgrad: bool) -> torch.Tensor:
        """fp32 master weight [O, I(, k, k)] -> the fp16 K-major operand of the forward GEMM ([O, taps * Ipad]) or of the
        data-gradient GEMM ([I, taps * Opad], transposed + tap-flipped) in ONE kernel pass over the master (torch's
        permute / flip / pad / cast chains of ops.pack_weight / train.dgrad_weight cost ~25 ms per step)."""
        w = self.p[name + ".weight"]
        O, I = w.shape[0], w.shape[1]
        taps = w.shape[2] * w.shape[3] if w.dim() == 4 else 1
        key = (name, "t" if dgrad else "w")
        buf = self._packed.get(key)
        if buf is None:
            rows, inner = (I, O) if dgrad else (O, I)
            buf = torch.zeros(rows, taps * ((inner + 63) // 64 * 64), device=self.dev, dtype=torch.float16)
            self._packed[key] = buf
        T.pack_master_weight(w, O, I, taps, buf, dgrad)
        return buf

    def weight(self, name: str, kind: int) -> torch.Tensor:
        key = (name, f"w{kind}")
        if key not in self._cache:
            if self.native_pack:
                self._cache[key] = self._pack_native(name, False)
            else:
                self._cache[key] = ops.pack_weight([(self.p[name + ".weight"], SEG_3x3 if kind == SEG_3x3_S2 else kind)])
        return self._cache[key]

    def weight_t(self, name: str) -> torch.Tensor:
        key = (name, "t")
        if key not in self._cache:
            self._cache[key] = self._pack_native(name, True) if self.native_pack else T.dgrad_weight(self.p[name + ".weight"])
        return self._cache[key]

    def bias(self, name: str) -> Optional[torch.Tensor]:
        return self.p.get(name + ".bias")

    def add_grad(self, name: str, g: torch.Tensor):
        self.g[name].add_(g.reshape(self.g[name].shape))

    def zero_grad(self):
        self.grad.zero_()


class Tape:
    def __init__(self, params: ParamSet, like: Optional["Tape"] = None, checkpointing: bool = False):
        self.P = params
        self.dev = params.dev
        self._bwd: List[Callable[[], None]] = []
        self.checkpointing = checkpointing
        if like is not None:          # a sub-tape of a checkpointed block: same workspaces
            self.partial, self.scratch = like.partial, like.scratch
        else:
            self.partial = torch.empty(16 << 20, device=self.dev, dtype=torch.float32)
            self.scratch = torch.empty(1 << 18, device=self.dev, dtype=torch.float32)

    def push(self, fn: Callable[[], None]):
        self._bwd.append(fn)

    def backward(self):
        for fn in reversed(self._bwd):
            fn()
        self._bwd.clear()

    def acc(self, t: TT, g: torch.Tensor):
        if not t.needs_grad:
            return
        if t.g is None:
            t.g = g
        else:
            o = torch.empty_like(t.g)
            ops.add_f16(None, t.g, g.contiguous(), o)
            t.g = o

    def new(self, rows: int, Cn: int) -> torch.Tensor:
        return torch.empty(rows, Cn, device=self.dev, dtype=torch.float16)


# ---------------------------------------------------------------------------------------------------------------------
# tape ops
# ---------------------------------------------------------------------------------------------------------------------
def _ld8(n: int) -> int:
    return (n + 7) // 8 * 8


def conv(tp: Tape, x: TT, name: str, k: int = 3, stride: int = 1, bias_tab: Optional[TT] = None,
         res: Optional[TT] = None) -> TT:
    """Conv2d(k in {1, 3}, padding k // 2, stride in {1, 2}) or Linear (k = 1) with the reference's parameter names
    `name`.weight / .bias.  bias_tab: a [>= B, N] matrix added per SAMPLE (the projected time embedding,
    models/unet_2d_blocks.py ResnetBlock2D) -- its gradient is the per-sample column sum.  res: fused residual add."""
    P = tp.P
    w = P.p[name + ".weight"]
    N, Cin = w.shape[0], w.shape[1]
    assert k == (w.shape[-1] if w.dim() == 4 else 1) and (stride == 1 or k == 3)
    B, H, W = x.B, x.H, x.W
    if stride == 2:
        H, W = H // 2, W // 2
    M = B * H * W if k == 3 else x.rows
    kind = SEG_1x1 if k == 1 else (SEG_3x3 if stride == 1 else SEG_3x3_S2)
    out = tp.new(M, _ld8(N))
    if out.shape[1] != N:
        out.zero_()
    bias = P.bias(name)
    bstride = 0
    if bias_tab is not None:
        bias = (bias_tab.v[:B].float() + (bias if bias is not None else 0.0)).contiguous()
        bstride = N
    ops.conv_gemm(None, [(x.v, Cin, kind)], P.weight(name, kind), out, M=M, N=N, B=B, H=H if k == 3 else 0,
                  W=W if k == 3 else 0, bias=bias, bias_bstride=bstride, res=res.v if res is not None else None,
                  partial=tp.partial)
    y = TT(out, B, H if k == 3 else x.H, W if k == 3 else x.W)

    def bwd():
        dy = y.g
        if dy is None:
            return
        if res is not None:
            tp.acc(res, dy)
        dyw, Hw, Ww = dy, H, W
        if stride == 2:                                     # zero insertion: the stride-1 kernels see a 2H x 2W problem
            dyw = T.scatter2x(dy, B, H, W)
            Hw, Ww = 2 * H, 2 * W
        want_b = P.bias(name) is not None and N % 8 == 0
        if P.native_pack and k == 3:
            # the tap-major fp32 result goes straight into the reference-layout flat gradient (no permute copy + add)
            _, db = T.conv_wgrad(x.v, Cin, dyw, N, B=B, H=Hw, W=Ww, taps=9, partial=tp.partial, want_bias=want_b,
                                 accumulate_into=P.g[name + ".weight"])
        else:
            dw, db = T.conv_wgrad(x.v, Cin, dyw, N, B=B, H=Hw if k == 3 else 0, W=Ww if k == 3 else 0, taps=k * k,
                                  partial=tp.partial, want_bias=want_b)
            P.add_grad(name + ".weight", dw if k == 3 else dw.reshape(N, Cin))
        if P.bias(name) is not None:
            P.add_grad(name + ".bias", db if want_b else dy[:, :N].float().sum(0))
        if bias_tab is not None:
            HW = H * W
            gt = torch.zeros(bias_tab.rows, N, device=tp.dev, dtype=torch.float16)
            gt[:B] = torch.stack([T.colsum(dy[b * HW:(b + 1) * HW], N) for b in range(B)], 0).half()
            tp.acc(bias_tab, gt)
        if x.needs_grad:
            dx = tp.new(x.rows, _ld8(Cin))
            if dx.shape[1] != Cin:
                dx.zero_()
            ops.conv_gemm(None, [(dyw, N, SEG_3x3 if k == 3 else SEG_1x1)], P.weight_t(name), dx, M=x.rows, N=Cin, B=B,
                          H=Hw if k == 3 else 0, W=Ww if k == 3 else 0, partial=tp.partial)
            tp.acc(x, dx if dx.shape[1] == x.C else dx[:, :x.C].contiguous())

    tp.push(bwd)
    return y


def gn(tp: Tape, x: TT, name: str, groups: int, eps: float, silu: bool) -> TT:
    gamma, beta = tp.P.p[name + ".weight"], tp.P.p[name + ".bias"]
    out = torch.empty_like(x.v)
    HW = x.H * x.W
    ops.groupnorm(None, x.v, x.C, None, 0, gamma, beta, out, tp.scratch, B=x.B, HW=HW, groups=groups, eps=eps, silu=silu)
    y = TT(out, x.B, x.H, x.W)

    def bwd():
        if y.g is None:
            return
        dx, dg, db = T.groupnorm_backward(x.v, y.g, gamma, beta, B=x.B, HW=HW, groups=groups, eps=eps, silu=silu)
        tp.P.add_grad(name + ".weight", dg)
        tp.P.add_grad(name + ".bias", db)
        tp.acc(x, dx)

    tp.push(bwd)
    return y


def ln(tp: Tape, x: TT, name: str) -> TT:
    gamma, beta = tp.P.p[name + ".weight"], tp.P.p[name + ".bias"]
    out = torch.empty_like(x.v)
    ops.layernorm(None, x.v, out, gamma, beta)
    y = TT(out, x.B, x.H, x.W)

    def bwd():
        if y.g is None:
            return
        dx, dg, db = T.layernorm_backward(x.v, y.g, gamma)
        tp.P.add_grad(name + ".weight", dg)
        tp.P.add_grad(name + ".bias", db)
        tp.acc(x, dx)

    tp.push(bwd)
    return y


def add(tp: Tape, a: TT, b: TT) -> TT:
    out = torch.empty_like(a.v)
    ops.add_f16(None, a.v, b.v, out)
    y = TT(out, a.B, a.H, a.W)

    def bwd():
        if y.g is not None:
            tp.acc(a, y.g)
            tp.acc(b, y.g)

    tp.push(bwd)
    return y


def cat(tp: Tape, a: TT, b: TT) -> TT:
    """torch.cat([a, b], dim=1) of the reference's up blocks (models/unet_2d_blocks.py:2559)."""
    y = TT(torch.cat([a.v, b.v], 1), a.B, a.H, a.W)

    def bwd():
        if y.g is not None:
            tp.acc(a, y.g[:, :a.C].contiguous())
            tp.acc(b, y.g[:, a.C:].contiguous())

    tp.push(bwd)
    return y


def replace_head(tp: Tape, src: TT, head: torch.Tensor, n_head: int, n_valid: int) -> TT:
    """cat(head[:, :n_head], src[:, n_head:n_valid]) -- `torch.cat((latents_mask, mask_pred), dim=1)` of the consistency
    pass (train/train.py:1393): the predicted attribute channels stay on the tape, the clean mask channels are data."""
    v = src.v.clone()
    v[:, :n_head] = head[:, :n_head]
    v[:, n_valid:] = 0
    y = TT(v, src.B, src.H, src.W)

    def bwd():
        if y.g is None:
            return
        g = y.g.clone()
        g[:, :n_head] = 0
        g[:, n_valid:] = 0
        tp.acc(src, g)

    tp.push(bwd)
    return y


def upsample(tp: Tape, x: TT) -> TT:
    out = tp.new(x.rows * 4, x.C)
    ops.upsample2x(None, x.v, out, B=x.B, H=x.H, W=x.W, Cn=x.C)
    y = TT(out, x.B, 2 * x.H, 2 * x.W)

    def bwd():
        if y.g is None:
            return
        tp.acc(x, T.pool2x2_sum(y.g, x.B, x.H, x.W))

    tp.push(bwd)
    return y


def silu(tp: Tape, x: TT) -> TT:
    y = TT(T.silu_f16(x.v), x.B, x.H, x.W)

    def bwd():
        if y.g is not None:
            tp.acc(x, T.silu_f16(x.v, y.g))

    tp.push(bwd)
    return y


def geglu_op(tp: Tape, proj: TT) -> TT:
    y = TT(T.geglu(proj.v), proj.B, proj.H, proj.W)

    def bwd():
        if y.g is not None:
            tp.acc(proj, T.geglu(proj.v, y.g))

    tp.push(bwd)
    return y


def _pack_rows(m: torch.Tensor) -> torch.Tensor:
    """[rows, d] column slice -> contiguous [rows, ceil(d/64)*64] (zero padded): the [N, K] operand of a GEMM."""
    rows, d = m.shape
    dst = torch.empty(rows, (d + 63) // 64 * 64, device=m.device, dtype=torch.float16)
    ops.to_nhwc(None, torch.as_strided(m, (1, d, rows, 1), (0, 1, m.stride(0), 1)), dst, dst.shape[1])
    return dst


def _pack_cols(m: torch.Tensor) -> torch.Tensor:
    """[rows, d] column slice -> its transpose, contiguous [d, ceil(rows/64)*64] (zero padded)."""
    rows, d = m.shape
    dst = torch.empty(d, (rows + 63) // 64 * 64, device=m.device, dtype=torch.float16)
    ops.to_nhwc(None, torch.as_strided(m, (1, rows, d, 1), (0, m.stride(0), 1, 1)), dst, dst.shape[1])
    return dst


def attention(tp: Tape, q: TT, k: TT, v: TT, heads: int, B: int) -> TT:
    """softmax(Q K^T / sqrt(d)) V per sample and head (Attention + AttnProcessor2_0 of the reference's transformer
    blocks).  q: [B*Nq, C]; k, v: [B*Nk, C]; any Nk (the score matrix is stored with a leading dimension padded to 8)."""
    Cn = q.C
    d = Cn // heads
    Nq, Nk = q.rows // B, k.rows // B
    Lp = _ld8(Nk)
    scale = d ** -0.5
    ao = tp.new(q.rows, Cn)
    # forward: the fused flash kernel of the inference path (one launch, nothing but O and the row log-sum-exps kept)
    flash_bwd = d % 8 == 0 and d <= 80
    lse2 = torch.empty(B * heads * Nq, device=tp.dev, dtype=torch.float32) if flash_bwd else None
    ops.attention(None, q.v, k.v, v.v, ao, B=B, heads=heads, Nq=Nq, Nk=Nk, d=d, scale=scale, lse2=lse2)
    y = TT(ao, q.B, q.H, q.W)

    def bwd():
        if y.g is None:
            return
        if flash_bwd:         # head dims <= 80: the flash backward kernel (csrc/attention_bwd_sm100.cu), one launch
            dq, dk, dv = T.attention_backward(q.v, k.v, v.v, ao, y.g, lse2, B=B, heads=heads, Nq=Nq, Nk=Nk, d=d, scale=scale)
            tp.acc(q, dq)
            tp.acc(k, dk)
            tp.acc(v, dv)
            return
        # wider heads (the 8x8 / 16x16 levels, few tokens): per (sample, head) the probabilities are recomputed in the
        # materialised form (S = Q K^T, row softmax) and the four gradient GEMMs run on the GEMM / wgrad kernels
        dq = torch.empty_like(q.v)            # every (sample, head) slice below is written in full
        dk = torch.empty_like(k.v)
        dv = torch.empty_like(v.v)
        for b in range(B):
            for h in range(heads):
                cs = slice(h * d, (h + 1) * d)
                rq, rk = slice(b * Nq, (b + 1) * Nq), slice(b * Nk, (b + 1) * Nk)
                do = y.g[rq, cs]
                p = (torch.zeros if Lp != Nk else torch.empty)(Nq, Lp, device=tp.dev, dtype=torch.float16)
                ops.conv_gemm(None, [(q.v[rq, cs], d, SEG_1x1)], _pack_rows(k.v[rk, cs]), p, M=Nq, N=Nk)      # S = Q K^T
                ops.softmax_rows(None, p, rows=Nq, n=Nk, scale=scale)                                         # P
                if v.needs_grad:
                    dvh, _ = T.conv_wgrad(do, d, p, Nk, B=1, H=0, W=0, taps=1, want_bias=False)               # dV = P^T dO
                    T.cvt_f32_f16(dvh, dv[rk, cs])
                dp = (torch.zeros if Lp != Nk else torch.empty)(Nq, Lp, device=tp.dev, dtype=torch.float16)
                ops.conv_gemm(None, [(do, d, SEG_1x1)], _pack_rows(v.v[rk, cs]), dp, M=Nq, N=Nk)              # dP = dO V^T
                T.softmax_backward(p, dp, Nk, scale)                                                          # dp := dS
                if q.needs_grad:
                    ops.conv_gemm(None, [(dp, Nk, SEG_1x1)], _pack_cols(k.v[rk, cs]), dq[rq, cs], M=Nq, N=d)  # dQ = dS K
                if k.needs_grad:
                    dkh, _ = T.conv_wgrad(q.v[rq, cs], d, dp, Nk, B=1, H=0, W=0, taps=1, want_bias=False)      # dK = dS^T Q
                    T.cvt_f32_f16(dkh, dk[rk, cs])
        tp.acc(q, dq)
        tp.acc(k, dk)
        tp.acc(v, dv)

    tp.push(bwd)
    return y


def checkpoint(tp: Tape, fn, *inputs: TT) -> TT:
    """Activation checkpointing of one block (`torch.utils.checkpoint` around each resnet / transformer when the
    reference's `gradient_checkpointing` is on, models/unet_2d_blocks.py:1172-1197): the forward keeps only the block's
    inputs and output; the backward re-runs the block on a private tape and unwinds that.  The kernels are
    deterministic, so the recomputed activations are bit-identical; the gradients differ only by the fp16 rounding of
    regrouped sums (a block input's gradient is summed inside the block before it joins the outer sum)."""
    if not tp.checkpointing:
        return fn(tp, *inputs)
    sub = Tape(tp.P, like=tp)
    out = fn(sub, *inputs)
    sub._bwd.clear()                                  # drops every intermediate activation of the block
    y = TT(out.v, out.B, out.H, out.W)

    def bwd():
        if y.g is None:
            return
        again = Tape(tp.P, like=tp)
        ins = [TT(i.v, i.B, i.H, i.W, i.needs_grad) for i in inputs]
        o = fn(again, *ins)
        o.g = y.g
        again.backward()
        for i, j in zip(inputs, ins):
            if j.g is not None:
                tp.acc(i, j.g)

    tp.push(bwd)
    return y


# ---------------------------------------------------------------------------------------------------------------------
# the networks (same walk as oracle/uni_oracle.py, which cites the reference line by line)
# ---------------------------------------------------------------------------------------------------------------------
def _sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, freq_shift=0), models/controlnet.py:282,909 (no parameters: host glue)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.reshape(-1, 1).float() * freqs[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def time_embedding(tp: Tape, net: str, cfg, t: torch.Tensor, B: int) -> TT:
    """silu(time_embedding(time_proj(t))) as a 128-row padded matrix (rows >= B carry no gradient); every resnet's
    time_emb_proj reads it."""
    e = torch.zeros(128, cfg.block_out_channels[0], device=tp.dev, dtype=torch.float16)
    e[:B] = _sinusoid(t.to(tp.dev).reshape(-1).expand(B) if t.numel() == 1 else t.to(tp.dev), cfg.block_out_channels[0]).half()
    x = TT(e, needs_grad=False)
    h = silu(tp, conv(tp, x, f"{net}.time_embedding.linear_1", k=1))
    return silu(tp, conv(tp, h, f"{net}.time_embedding.linear_2", k=1))


def resnet(tp: Tape, p: str, x: TT, temb_act: TT, cfg) -> TT:
    h = gn(tp, x, p + ".norm1", cfg.norm_num_groups, cfg.norm_eps, True)
    tproj = conv(tp, temb_act, p + ".time_emb_proj", k=1)
    h = conv(tp, h, p + ".conv1", k=3, bias_tab=tproj)
    h = gn(tp, h, p + ".norm2", cfg.norm_num_groups, cfg.norm_eps, True)
    sc = conv(tp, x, p + ".conv_shortcut", k=1) if (p + ".conv_shortcut.weight") in tp.P.p else x
    return conv(tp, h, p + ".conv2", k=3, res=sc)


def transformer_2d(tp: Tape, p: str, x: TT, ctx: TT, cfg) -> TT:
    B, heads = x.B, cfg.num_heads
    h = gn(tp, x, p + ".norm", cfg.norm_num_groups, 1e-6, False)
    h = conv(tp, h, p + ".proj_in", k=1)
    t = p + ".transformer_blocks.0"
    y = ln(tp, h, t + ".norm1")
    a = attention(tp, conv(tp, y, t + ".attn1.to_q", k=1), conv(tp, y, t + ".attn1.to_k", k=1),
                  conv(tp, y, t + ".attn1.to_v", k=1), heads, B)
    h = conv(tp, a, t + ".attn1.to_out.0", k=1, res=h)
    y = ln(tp, h, t + ".norm2")
    a = attention(tp, conv(tp, y, t + ".attn2.to_q", k=1), conv(tp, ctx, t + ".attn2.to_k", k=1),
                  conv(tp, ctx, t + ".attn2.to_v", k=1), heads, B)
    h = conv(tp, a, t + ".attn2.to_out.0", k=1, res=h)
    y = ln(tp, h, t + ".norm3")
    f = geglu_op(tp, conv(tp, y, t + ".ff.net.0.proj", k=1))
    h = conv(tp, f, t + ".ff.net.2", k=1, res=h)
    return conv(tp, h, p + ".proj_out", k=1, res=x)


def _resnet_ck(tp: Tape, p: str, x: TT, temb_act: TT, cfg) -> TT:
    return checkpoint(tp, lambda t, a, e: resnet(t, p, a, e, cfg), x, temb_act)


def _transformer_ck(tp: Tape, p: str, x: TT, ctx: TT, cfg) -> TT:
    return checkpoint(tp, lambda t, a, c: transformer_2d(t, p, a, c, cfg), x, ctx)


def down_blocks(tp: Tape, net: str, cfg, sample: TT, temb: TT, ctx: TT):
    skips = [sample]
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            sample = _resnet_ck(tp, f"{net}.down_blocks.{i}.resnets.{j}", sample, temb, cfg)
            if cfg.down_has_attn[i]:
                sample = _transformer_ck(tp, f"{net}.down_blocks.{i}.attentions.{j}", sample, ctx, cfg)
            skips.append(sample)
        if i != nb - 1:
            sample = conv(tp, sample, f"{net}.down_blocks.{i}.downsamplers.0.conv", k=3, stride=2)
            skips.append(sample)
    return sample, skips


def mid_block(tp: Tape, net: str, cfg, sample: TT, temb: TT, ctx: TT) -> TT:
    sample = _resnet_ck(tp, f"{net}.mid_block.resnets.0", sample, temb, cfg)
    sample = _transformer_ck(tp, f"{net}.mid_block.attentions.0", sample, ctx, cfg)
    return _resnet_ck(tp, f"{net}.mid_block.resnets.1", sample, temb, cfg)


def up_blocks(tp: Tape, net: str, cfg, sample: TT, skips: Sequence[TT], temb: TT, ctx: TT) -> TT:
    skips = list(skips)
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            sample = cat(tp, sample, skips.pop())
            sample = _resnet_ck(tp, f"{net}.up_blocks.{i}.resnets.{j}", sample, temb, cfg)
            if cfg.up_has_attn[i]:
                sample = _transformer_ck(tp, f"{net}.up_blocks.{i}.attentions.{j}", sample, ctx, cfg)
        if i != nb - 1:
            sample = conv(tp, upsample(tp, sample), f"{net}.up_blocks.{i}.upsamplers.0.conv", k=3)
    return sample


def out_head(tp: Tape, net: str, cfg, sample: TT) -> TT:
    h = gn(tp, sample, f"{net}.conv_norm_out", cfg.norm_num_groups, cfg.norm_eps, True)
    return conv(tp, h, f"{net}.conv_out", k=3)


def unet_forward(tp: Tape, net: str, cfg, sample: TT, t: torch.Tensor, ctx: TT, down_add: Optional[Sequence[TT]] = None,
                 mid_add: Optional[TT] = None):
    """UNet2DConditionModel.forward (models/controlnet.py:781-1166): (prediction, raw_down[12], raw_mid)."""
    temb = time_embedding(tp, net, cfg, t, sample.B)
    h = conv(tp, sample, f"{net}.conv_in", k=3)
    h, skips = down_blocks(tp, net, cfg, h, temb, ctx)
    raw_down = list(skips)
    if down_add is not None:
        skips = [add(tp, s, r) for s, r in zip(skips, down_add)]
    h = mid_block(tp, net, cfg, h, temb, ctx)
    raw_mid = h
    if mid_add is not None:
        h = add(tp, h, mid_add)
    h = up_blocks(tp, net, cfg, h, skips, temb, ctx)
    return out_head(tp, net, cfg, h), raw_down, raw_mid


def attr_encoder_forward(tp: Tape, net: str, cfg, t: torch.Tensor, ctx: TT, cond: TT):
    """AttributeEncoderModel.forward (models/controlnet.py:1657-1778): (zero-conv'd down[12], zero-conv'd mid, raw_down, raw_mid)."""
    temb = time_embedding(tp, net, cfg, t, cond.B)
    h = conv(tp, cond, f"{net}.conv_in", k=3)
    h, skips = down_blocks(tp, net, cfg, h, temb, ctx)
    h = mid_block(tp, net, cfg, h, temb, ctx)
    down = [conv(tp, s, f"{net}.controlnet_down_blocks.{i}", k=1) for i, s in enumerate(skips)]
    mid = conv(tp, h, f"{net}.controlnet_mid_block", k=1)
    return down, mid, skips, h


def attr_decoder_forward(tp: Tape, net: str, cfg, sample: TT, down_res: Sequence[TT], t: torch.Tensor, ctx: TT,
                         down_add: Sequence[TT], mid_add: TT) -> TT:
    """AttributeDecoderModel.forward (models/controlnet.py:2342-2527)."""
    temb = time_embedding(tp, net, cfg, t, sample.B)
    skips = [conv(tp, r, f"{net}.control_down_blocks.{i}", k=1, res=s) for i, (s, r) in enumerate(zip(down_res, down_add))]
    sample = conv(tp, mid_add, f"{net}.control_mid_block", k=1, res=sample)
    h = up_blocks(tp, net, cfg, sample, skips, temb, ctx)
    return out_head(tp, net, cfg, h)


# ---------------------------------------------------------------------------------------------------------------------
# the training step
# ---------------------------------------------------------------------------------------------------------------------
def to_matrix(x: torch.Tensor, dev, needs_grad: bool = False) -> TT:
    """NCHW (or [B, L, D] tokens) fp32 / fp16 -> NHWC fp16 matrix with the channel count padded to a multiple of 8."""
    if x.dim() == 3:
        B, Ln, D = x.shape
        return TT(x.to(dev).reshape(B * Ln, D).half().contiguous(), B, Ln, 1, needs_grad)
    B, Cn, H, W = x.shape
    m = torch.zeros(B * H * W, _ld8(Cn), device=dev, dtype=torch.float16)
    m[:, :Cn] = x.to(dev).permute(0, 2, 3, 1).reshape(B * H * W, Cn).half()
    return TT(m, B, H, W, needs_grad)


def to_nchw(t: TT, Cn: int) -> torch.Tensor:
    return t.v[:, :Cn].float().reshape(t.B, t.H, t.W, Cn).permute(0, 3, 1, 2).contiguous()


def reference_losses(img_pred, mask_pred_full, img_target, attr_target):
    """train/train.py:1350-1373: MSE on the RGB prediction, 10 x MSE on the 24 attribute channels (the clean mask channels
    are cut off first), 0.01 x the cosine contrastive term between the first two samples of the batch."""
    import torch.nn.functional as F
    mask_pred = mask_pred_full[:, 4:]
    material, albedo, spec = mask_pred[:, :4], mask_pred[:, 8:12], mask_pred[:, 12:16]
    temperature = 0.1
    loss = F.mse_loss(img_pred.float(), img_target.float()) + 10.0 * F.mse_loss(mask_pred.float(), attr_target.float())
    if mask_pred.shape[0] >= 2:
        m_dis = F.cosine_similarity(material[0].reshape(-1).float(), material[1].reshape(-1).float(), dim=0) / temperature
        a_dis = F.cosine_similarity(albedo[0].reshape(-1).float(), albedo[1].reshape(-1).float(), dim=0) / temperature
        s_dis = F.cosine_similarity(spec[0].reshape(-1).float(), spec[1].reshape(-1).float(), dim=0) / temperature
        pos = torch.exp(a_dis)
        loss = loss - 0.01 * torch.log(pos / (pos + torch.exp(m_dis) + torch.exp(s_dis)))
    return loss


def reference_losses_inverse(img_pred, mask_pred_full, img_pred_c, img_target, attr_target):
    """The `is_inv_rendering` branch (train/train.py:1375-1416): loss_img + loss_mask + 0.8 x the consistency loss of the
    second pass (the contrastive term and the x10 mask weight are dropped there)."""
    import torch.nn.functional as F
    return (F.mse_loss(img_pred.float(), img_target.float()) + F.mse_loss(mask_pred_full[:, 4:].float(), attr_target.float()) +
            0.8 * F.mse_loss(img_pred_c.float(), img_target.float()))


def allreduce_gradients(flat_grad: torch.Tensor, bucket_bytes: int = 64 << 20, group=None) -> int:
    """Average the flat gradient buffer over the data-parallel ranks in fixed-size buckets (one collective per bucket, so a
    caller can overlap them with the tail of its backward).  Returns the number of collectives issued."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    step = max(1, bucket_bytes // flat_grad.element_size())
    n = 0
    for off in range(0, flat_grad.numel(), step):
        chunk = flat_grad[off:off + step]
        dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=group)
        chunk.div_(world)
        n += 1
    return n


class DualStreamTrainer:
    """One optimizer step of the reference's training loop (train/train.py:1324-1427) on the B200 kernels.

    nets = {"unet": sd, "enc": sd, "dec": sd} with the reference's parameter names; cfgs the matching NetConfig-like
    objects (block_out_channels, layers_per_block, num_heads, norm_num_groups, norm_eps, down_has_attn, up_has_attn)."""

    def __init__(self, nets: Dict[str, SD], cfgs: Dict[str, object], *, lr: float = 1e-5, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: Optional[float] = 1.0, loss_scale: float = 1024.0,
                 gradient_checkpointing: bool = False, use_cuda_graph: bool = False, device="cuda"):
        self.P = ParamSet(nets, device)
        self.gradient_checkpointing = gradient_checkpointing      # the reference's enable_gradient_checkpointing()
        # use_cuda_graph: `step` captures forward + loss + backward (every kernel launch, the weight packing and the
        # loss-head autograd) into ONE CUDA graph on its first call and replays it afterwards: the ~13 k launches of a
        # step then cost no host time.  Shapes (and whether the consistency pass runs) are fixed by the first batch.
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.cfgs = cfgs
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.max_grad_norm, self.loss_scale = max_grad_norm, loss_scale
        self.dev = self.P.dev

    # -- forward + backward ------------------------------------------------------------------------------------------
    def forward_backward(self, x_img, t_img, x_attr, t_attr, ehs, img_target, attr_target, loss_fn=None, cycle=None):
        """The 3-call forward, the loss, and the whole backward: gradients accumulate (scaled by loss_scale) in P.grad.
        x_img [B,4,H,W] noisy RGB latents; x_attr [B,28,H,W] = cat(mask latents, noisy attribute latents); ehs [B,L,D].
        cycle = (x_img_c, t_img_c) adds the consistency pass of inverse-rendering batches (train/train.py:1375-1416): the
        encoder and the UNet run again on cat(clean mask latents, PREDICTED attributes) at attribute timestep 0 -- the
        gradient of that pass flows back through the prediction into the first pass -- and the loss becomes
        reference_losses_inverse.  Returns (loss, img_pred [B,4,H,W] fp32, mask_pred [B,28,H,W] fp32)."""
        tp = Tape(self.P, checkpointing=self.gradient_checkpointing)
        dev = self.dev
        ctx = to_matrix(ehs, dev)
        xi, xa = to_matrix(x_img, dev), to_matrix(x_attr, dev)
        down, mid, raw_a, raw_a_mid = attr_encoder_forward(tp, "enc", self.cfgs["enc"], t_attr, ctx, xa)
        img, raw_u, raw_u_mid = unet_forward(tp, "unet", self.cfgs["unet"], xi, t_img, ctx, down, mid)
        msk = attr_decoder_forward(tp, "dec", self.cfgs["dec"], raw_a_mid, raw_a, t_attr, ctx, raw_u, raw_u_mid)
        c_img, c_msk = self.P.p["unet.conv_out.weight"].shape[0], self.P.p["dec.conv_out.weight"].shape[0]
        heads = [(img, c_img), (msk, c_msk)]
        if cycle is not None:
            x_img_c, t_img_c = cycle
            cond2 = replace_head(tp, msk, xa.v, 4, c_msk)
            t0 = torch.zeros(xi.B, device=dev)
            down2, mid2, _, _ = attr_encoder_forward(tp, "enc", self.cfgs["enc"], t0, ctx, cond2)
            img_c, _, _ = unet_forward(tp, "unet", self.cfgs["unet"], to_matrix(x_img_c, dev), t_img_c, ctx, down2, mid2)
            heads.append((img_c, c_img))
        preds = [to_nchw(t, cn).requires_grad_(True) for t, cn in heads]
        if cycle is not None:
            loss = (loss_fn or reference_losses_inverse)(preds[0], preds[1], preds[2], img_target.to(dev), attr_target.to(dev))
        else:
            loss = (loss_fn or reference_losses)(preds[0], preds[1], img_target.to(dev), attr_target.to(dev))
        (loss * self.loss_scale).backward()
        # seed the tape: later heads first is irrelevant -- every head's gradient is set before the tape unwinds
        for (t, cn), pr in zip(heads, preds):
            g = torch.zeros_like(t.v)
            g[:, :cn] = pr.grad.permute(0, 2, 3, 1).reshape(-1, cn).half()
            tp.acc(t, g)
        tp.backward()
        return loss.detach(), preds[0].detach(), preds[1].detach()

    # -- optimizer ---------------------------------------------------------------------------------------------------
    def optimizer_step(self) -> Dict[str, float]:
        """All-reduce (if a process group is up), overflow check, global-norm clip, AdamW, zero_grad."""
        P = self.P
        n_coll = allreduce_gradients(P.grad)
        inv = 1.0 / self.loss_scale
        gnorm = float(torch.linalg.vector_norm(P.grad).item()) * inv
        info = {"grad_norm": gnorm, "collectives": n_coll, "skipped": 0.0}
        if not math.isfinite(gnorm):             # fp16 overflow somewhere in the backward: skip the step (GradScaler semantics)
            info["skipped"] = 1.0
            P.zero_grad()
            return info
        clip = 1.0
        if self.max_grad_norm is not None and gnorm > self.max_grad_norm:
            clip = self.max_grad_norm / (gnorm + 1e-6)
        if P.m is None:
            P.m, P.v = torch.zeros_like(P.flat), torch.zeros_like(P.flat)
        P.steps += 1
        T.adamw_step(P.flat, P.grad, P.m, P.v, lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.wd,
                     step=P.steps, grad_scale=inv * clip)
        P.invalidate()
        P.zero_grad()
        return info

    def step(self, *batch, **kw):
        if self.use_cuda_graph:
            loss = self._graph_step(batch, kw)
        else:
            loss, _, _ = self.forward_backward(*batch, **kw)
        info = self.optimizer_step()
        info["loss"] = float(loss.item())
        return info

    def _graph_step(self, batch, kw):
        dev = self.dev
        cyc = kw.get("cycle")
        flat = list(batch) + (list(cyc) if cyc is not None else [])
        if self._graph is None:
            self._static = [t.detach().to(dev).clone() for t in flat]
            nb = len(batch)

            def run():
                st = self._static
                k2 = dict(kw)
                if cyc is not None:
                    k2["cycle"] = (st[nb], st[nb + 1])
                return self.forward_backward(*st[:nb], **k2)

            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):               # eager warm-up: lazy kernel attributes, allocator, autograd
                run()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.P.zero_grad()
            self.P.invalidate()                         # the weight packing must be part of the captured work
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._graph_out = run()
            self.P.invalidate()
        for dst, src in zip(self._static, flat):
            if tuple(dst.shape) != tuple(src.shape):
                raise ValueError("use_cuda_graph: batch shapes are fixed by the first step")
            dst.copy_(src)
        self._graph.replay()
        return self._graph_out[0]
