"""Training slice (first row of the backward): forward AND backward of a ResnetBlock2D on the
B200 kernels -- the block the three networks are mostly made of (22 per stream) and that activation checkpointing
re-runs (models/unet_2d_blocks.py:1172-1197); train/train.py:1324-1427 runs the modules forward, takes an MSE loss and
calls `accelerator.backward`.

What maps to what (per conv / norm of the block):
  dX  of a conv   = the SAME implicit-GEMM forward kernel (unib200_conv_gemm) on dY with the weights repacked
                    (3x3 taps flipped, in / out channels swapped) -- `dgrad_weight`
  dW  of a conv   = unib200_conv_wgrad: pixels are the contraction dimension, both operands read MN-major from their NHWC
                    tensors by tcgen05 (csrc/wgrad_sm100.cu);  db = column sums of dY
  GroupNorm+SiLU  = unib200_groupnorm_backward (dx, dgamma, dbeta)
  time embedding  = d(time_emb_proj output)[b, c] = sum over pixels of dh1 -- the same column-sum kernel per sample

Scope of this slice: ResnetBlock2D (with or without the 1x1 shortcut) in fp16 storage / fp32 accumulation, gradients
w.r.t. the input, the projected time embedding and every parameter, checked against torch autograd of the oracle
(tests/test_train_gpu.py).  NOT yet built: attention / LayerNorm / GEGLU backward, the optimizer step, bucketed gradient
all-reduce -- the rest of row 8f-3.  No CPU fallback: everything below launches libunib200.so kernels."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from . import ops
from .ops import SEG_1x1, SEG_3x3


def dgrad_weight(w: torch.Tensor) -> torch.Tensor:
    """Packed weights of the data-gradient convolution: dX = conv(dY, W') with W'[ci, co, ky, kx] = W[co, ci, 2-ky, 2-kx]
    (3x3, pad 1, stride 1) or W'[ci, co] = W[co, ci] (1x1 / linear)."""
    w = w.detach().float()
    if w.dim() == 2:
        w = w[:, :, None, None]
    wt = w.permute(1, 0, 2, 3).flip(2, 3).contiguous()
    return ops.pack_weight([(wt, SEG_3x3 if w.shape[-1] == 3 else SEG_1x1)])


def pack_master_weight(w: torch.Tensor, O: int, I: int, taps: int, out: torch.Tensor, dgrad: bool) -> None:
    """unib200_pack_master_weight: fp32 master weight [O, I, taps] -> the valid columns of the fp16 operand buffer `out`
    ([O, taps * Ipad] forward, [I, taps * Opad] data gradient; padding columns were zeroed at allocation)."""
    assert w.dtype == torch.float32 and w.is_contiguous() and out.dtype == torch.float16 and out.is_contiguous()
    L.check(L.load().unib200_pack_master_weight(None, w.data_ptr(), O, I, taps, out.data_ptr(), int(dgrad),
                                                torch.cuda.current_stream().cuda_stream), "pack_master_weight")


def conv_wgrad(x: torch.Tensor, C_in: int, dy: torch.Tensor, N: int, *, B: int, H: int, W: int, taps: int,
               partial: Optional[torch.Tensor] = None, want_bias: bool = True,
               accumulate_into: Optional[torch.Tensor] = None):
    """(dW fp32 [N, C_in, k, k], db fp32 [N] | None) of a stride-1 conv (taps 9: 3x3 pad 1; taps 1: 1x1 / linear with
    H = W = 0) from its forward input x [M, >= C_in] and the output gradient dy [M, >= N], both fp16 NHWC matrices."""
    lib = L.load()
    M = x.shape[0]
    dw = torch.empty(N, taps, C_in, device=x.device, dtype=torch.float32)
    db = torch.empty(N, device=x.device, dtype=torch.float32) if want_bias else None
    d = L.WgradDesc()
    d.x, d.C, d.ldx = x.data_ptr(), C_in, x.stride(0)
    d.dy, d.N, d.lddy = dy.data_ptr(), N, dy.stride(0)
    d.M, d.B, d.H, d.W, d.taps = M, B, H, W, taps
    d.dw, d.db = dw.data_ptr(), db.data_ptr() if db is not None else None
    if partial is not None:
        d.partial, d.partial_bytes = partial.data_ptr(), partial.numel() * 4
    L.check(lib.unib200_conv_wgrad(None, C.byref(d), torch.cuda.current_stream().cuda_stream), "conv_wgrad")
    if accumulate_into is not None:
        # grad[n][c][t] += dw[n][t][c]: added into the reference-layout fp32 gradient by one kernel; returns (None, db)
        g = accumulate_into
        assert g.dtype == torch.float32 and g.is_contiguous() and g.numel() == N * taps * C_in
        L.check(lib.unib200_wgrad_scatter_add(None, dw.data_ptr(), g.data_ptr(), N, taps, C_in,
                                              torch.cuda.current_stream().cuda_stream), "wgrad_scatter_add")
        return None, db
    k = 3 if taps == 9 else 1
    return dw.reshape(N, k, k, C_in).permute(0, 3, 1, 2).contiguous(), db


def colsum(x: torch.Tensor, N: int) -> torch.Tensor:
    """fp32 [N] column sums of the fp16 matrix x [M, >= N] (bias gradients; per-sample time-embedding gradients)."""
    out = torch.empty(N, device=x.device, dtype=torch.float32)
    L.check(L.load().unib200_colsum(None, x.data_ptr(), x.stride(0), x.shape[0], N, out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "colsum")
    return out


def groupnorm_backward(x: torch.Tensor, dz: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, B: int, HW: int,
                       groups: int, eps: float, silu: bool):
    """(dx fp16 [B*HW, C], dgamma fp32 [C], dbeta fp32 [C]) of z = silu(group_norm(x)) (or group_norm(x)) given dz."""
    lib = L.load()
    Cn = x.shape[1]
    dx = torch.empty(B * HW, Cn, device=x.device, dtype=torch.float16)
    dg = torch.empty(Cn, device=x.device, dtype=torch.float32)
    dbt = torch.empty(Cn, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * B * Cn, device=x.device, dtype=torch.float32)
    d = L.GnBwdDesc()
    d.x, d.ldx, d.dz, d.ldz, d.dx, d.lddx = x.data_ptr(), x.stride(0), dz.data_ptr(), dz.stride(0), dx.data_ptr(), Cn
    d.gamma, d.beta, d.dgamma, d.dbeta, d.scratch = gamma.data_ptr(), beta.data_ptr(), dg.data_ptr(), dbt.data_ptr(), scratch.data_ptr()
    d.B, d.HW, d.C, d.groups, d.silu, d.eps = B, HW, Cn, groups, int(silu), eps
    L.check(lib.unib200_groupnorm_backward(None, C.byref(d), torch.cuda.current_stream().cuda_stream), "groupnorm_backward")
    return dx, dg, dbt


class ResnetBlockTrainer:
    """ResnetBlock2D(in, out, temb_channels, groups, eps, "default", "silu") forward + backward on the kernels.

    `sd` holds the block's parameters under diffusers names (norm1, conv1, time_emb_proj, norm2, conv2[, conv_shortcut]).
    Activations are NHWC fp16 matrices [B*H*W, C]; `temb_proj` is the already projected time embedding
    time_emb_proj(silu(temb)) as fp32 [B, out] (its Linear belongs to the embedding MLP, not to this slice)."""

    def __init__(self, sd: Dict[str, torch.Tensor], groups: int = 32, eps: float = 1e-5, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the training slice runs on CUDA (sm_100a) only; there is no CPU fallback")
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()       # noqa: E731
        self.dev, self.groups, self.eps = dev, groups, eps
        self.cin, self.cout = sd["conv1.weight"].shape[1], sd["conv1.weight"].shape[0]
        self.g1, self.b1 = f32(sd["norm1.weight"]), f32(sd["norm1.bias"])
        self.g2, self.b2 = f32(sd["norm2.weight"]), f32(sd["norm2.bias"])
        self.w1 = ops.pack_weight([(sd["conv1.weight"], SEG_3x3)]).to(dev)
        self.w2 = ops.pack_weight([(sd["conv2.weight"], SEG_3x3)]).to(dev)
        self.w1_t = dgrad_weight(sd["conv1.weight"]).to(dev)
        self.w2_t = dgrad_weight(sd["conv2.weight"]).to(dev)
        self.bias1, self.bias2 = f32(sd["conv1.bias"]), f32(sd["conv2.bias"])
        self.has_sc = "conv_shortcut.weight" in sd
        if self.has_sc:
            self.wsc = ops.pack_weight([(sd["conv_shortcut.weight"], SEG_1x1)]).to(dev)
            self.wsc_t = dgrad_weight(sd["conv_shortcut.weight"]).to(dev)
            self.bias_sc = f32(sd["conv_shortcut.bias"])
        self.scratch = torch.empty(1 << 18, device=dev, dtype=torch.float32)
        self.partial = torch.empty(16 << 20, device=dev, dtype=torch.float32)
        self.saved = None

    def _gn(self, x, gamma, beta, B, HW):
        out = torch.empty_like(x)
        ops.groupnorm(None, x, x.shape[1], None, 0, gamma, beta, out, self.scratch, B=B, HW=HW, groups=self.groups,
                      eps=self.eps, silu=True)
        return out

    def forward(self, x: torch.Tensor, temb_proj: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        """x: fp16 [B*H*W, cin]; temb_proj: fp32 [B, cout].  Saves what the backward needs (x, n1, h1, n2)."""
        M = B * H * W
        n1 = self._gn(x, self.g1, self.b1, B, H * W)
        h1 = torch.empty(M, self.cout, device=self.dev, dtype=torch.float16)
        bias_tab = (temb_proj + self.bias1[None]).contiguous()                 # conv1 bias + time embedding, per sample
        ops.conv_gemm(None, [(n1, self.cin, SEG_3x3)], self.w1, h1, M=M, N=self.cout, B=B, H=H, W=W, bias=bias_tab,
                      bias_bstride=self.cout)
        n2 = self._gn(h1, self.g2, self.b2, B, H * W)
        out = torch.empty(M, self.cout, device=self.dev, dtype=torch.float16)
        if self.has_sc:
            wcat = torch.cat([self.w2, self.wsc], 1).contiguous()              # shortcut accumulated into conv2's tile
            ops.conv_gemm(None, [(n2, self.cout, SEG_3x3), (x, self.cin, SEG_1x1)], wcat, out, M=M, N=self.cout, B=B, H=H,
                          W=W, bias=self.bias2 + self.bias_sc)
        else:
            ops.conv_gemm(None, [(n2, self.cout, SEG_3x3)], self.w2, out, M=M, N=self.cout, B=B, H=H, W=W, bias=self.bias2,
                          res=x)
        self.saved = (x, n1, h1, n2, B, H, W)
        return out

    def backward(self, dout: torch.Tensor) -> Dict[str, torch.Tensor]:
        """dout: fp16 [B*H*W, cout].  Returns the gradients: "x" (fp16 NHWC), "temb_proj" (fp32 [B, cout]) and one fp32
        tensor per parameter under its diffusers name."""
        x, n1, h1, n2, B, H, W = self.saved
        M, HW = B * H * W, H * W
        g: Dict[str, torch.Tensor] = {}
        # conv2 (+ shortcut): weight / bias gradients, data gradient through the flipped-weight forward kernel
        g["conv2.weight"], g["conv2.bias"] = conv_wgrad(n2, self.cout, dout, self.cout, B=B, H=H, W=W, taps=9,
                                                        partial=self.partial)
        dn2 = torch.empty(M, self.cout, device=self.dev, dtype=torch.float16)
        ops.conv_gemm(None, [(dout, self.cout, SEG_3x3)], self.w2_t, dn2, M=M, N=self.cout, B=B, H=H, W=W)
        dx_sc = dout
        if self.has_sc:
            g["conv_shortcut.weight"], g["conv_shortcut.bias"] = conv_wgrad(x, self.cin, dout, self.cout, B=B, H=0, W=0,
                                                                            taps=1, partial=self.partial)
            dx_sc = torch.empty(M, self.cin, device=self.dev, dtype=torch.float16)
            ops.conv_gemm(None, [(dout, self.cout, SEG_1x1)], self.wsc_t, dx_sc, M=M, N=self.cin, B=B)
        # SiLU + GroupNorm 2
        dh1, g["norm2.weight"], g["norm2.bias"] = groupnorm_backward(h1, dn2, self.g2, self.b2, B=B, HW=HW,
                                                                     groups=self.groups, eps=self.eps, silu=True)
        # conv1: dW, db; the time-embedding gradient is the per-sample column sum of dh1
        g["conv1.weight"], g["conv1.bias"] = conv_wgrad(n1, self.cin, dh1, self.cout, B=B, H=H, W=W, taps=9,
                                                        partial=self.partial)
        g["temb_proj"] = torch.stack([colsum(dh1[b * HW:(b + 1) * HW], self.cout) for b in range(B)], 0)
        dn1 = torch.empty(M, self.cin, device=self.dev, dtype=torch.float16)
        ops.conv_gemm(None, [(dh1, self.cout, SEG_3x3)], self.w1_t, dn1, M=M, N=self.cin, B=B, H=H, W=W)
        # SiLU + GroupNorm 1, plus the shortcut / identity branch
        dx, g["norm1.weight"], g["norm1.bias"] = groupnorm_backward(x, dn1, self.g1, self.b1, B=B, HW=HW,
                                                                    groups=self.groups, eps=self.eps, silu=True)
        gx = torch.empty_like(dx)
        ops.add_f16(None, dx, dx_sc.contiguous(), gx)
        g["x"] = gx
        return g


# ---------------------------------------------------------------------------------------------------------------------
# BasicTransformerBlock
# ---------------------------------------------------------------------------------------------------------------------
def _stream():
    return torch.cuda.current_stream().cuda_stream


def layernorm_backward(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5):
    """(dx fp16 [rows, C], dgamma fp32 [C], dbeta fp32 [C]) of y = layer_norm(x) * gamma + beta given dy."""
    rows, Cn = x.shape
    dx = torch.empty_like(x)
    gb = torch.empty(2 * Cn, device=x.device, dtype=torch.float32)
    scratch = torch.empty(592 * 2 * Cn, device=x.device, dtype=torch.float32)
    L.check(L.load().unib200_layernorm_backward(None, x.data_ptr(), dy.data_ptr(), dx.data_ptr(), gamma.data_ptr(),
                                                gb.data_ptr(), scratch.data_ptr(), scratch.numel(), rows, Cn, eps, _stream()),
            "layernorm_backward")
    return dx, gb[:Cn], gb[Cn:]


def geglu(proj: torch.Tensor, dout: Optional[torch.Tensor] = None) -> torch.Tensor:
    """forward (dout None): a * gelu(g) of proj = [a | g]; backward: dproj given dout."""
    rows, two = proj.shape
    inner = two // 2
    out = torch.empty(rows, inner if dout is None else two, device=proj.device, dtype=torch.float16)
    L.check(L.load().unib200_geglu(None, proj.data_ptr(), dout.data_ptr() if dout is not None else None, out.data_ptr(),
                                   rows, inner, _stream()), "geglu")
    return out


def softmax_backward(p: torch.Tensor, dp: torch.Tensor, n: int, scale: float):
    """In place on dp: dS = scale * P o (dP - rowsum(dP o P)) over the first n columns of the fp16 matrices p / dp."""
    L.check(L.load().unib200_softmax_backward(None, p.data_ptr(), dp.data_ptr(), p.shape[0], n, dp.stride(0), float(scale),
                                              _stream()), "softmax_backward")


def cvt_f32_f16(src: torch.Tensor, dst: torch.Tensor):
    """fp32 [rows, cols] contiguous -> the fp16 matrix view dst [rows, cols] (any leading dimension)."""
    rows, cols = dst.shape
    L.check(L.load().unib200_cvt_f32_f16(None, src.data_ptr(), dst.data_ptr(), rows, cols, dst.stride(0), _stream()), "cvt_f32_f16")


def silu_f16(x: torch.Tensor, dy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """silu(x) (dy None) or dy * silu'(x), elementwise on fp16."""
    out = torch.empty_like(x)
    L.check(L.load().unib200_silu_f16(None, x.data_ptr(), dy.data_ptr() if dy is not None else None, out.data_ptr(),
                                      x.numel(), _stream()), "silu_f16")
    return out


def scatter2x(x: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """[B*H*W, C] -> [B*2H*2W, C] with x at the even pixels and zeros elsewhere (stride-2 conv gradients)."""
    out = torch.empty(4 * x.shape[0], x.shape[1], device=x.device, dtype=torch.float16)
    L.check(L.load().unib200_scatter2x(None, x.data_ptr(), out.data_ptr(), B, H, W, x.shape[1], _stream()), "scatter2x")
    return out


def pool2x2_sum(x: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """[B*2H*2W, C] -> [B*H*W, C]: 2x2 block sums (adjoint of nearest-2x upsampling)."""
    out = torch.empty(x.shape[0] // 4, x.shape[1], device=x.device, dtype=torch.float16)
    L.check(L.load().unib200_pool2x2_sum(None, x.data_ptr(), out.data_ptr(), B, H, W, x.shape[1], _stream()), "pool2x2_sum")
    return out


def attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, dout: torch.Tensor,
                       lse2: torch.Tensor, *, B: int, heads: int, Nq: int, Nk: int, d: int, scale: float):
    """(dq, dk, dv) fp16, shaped like q / k / v, of O = softmax(Q K^T scale) V by the flash backward kernel (head dims
    <= 80): q / k / v / o / dout are fp16 matrices [B*N, >= heads*d] (head h in columns [h*d, (h+1)*d)), lse2 the
    forward's log-sum-exp output (ops.attention(..., lse2=))."""
    dev = q.device
    C_ = heads * d
    dq_acc = torch.zeros(B * Nq, C_, device=dev, dtype=torch.float32)
    dk = torch.empty(B * Nk, C_, device=dev, dtype=torch.float16)
    dv = torch.empty(B * Nk, C_, device=dev, dtype=torch.float16)
    D = torch.empty(B * heads * Nq, device=dev, dtype=torch.float32)
    a = L.AttnBwdDesc()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0)
    a.o, a.ldo, a.dout, a.lddo = o.data_ptr(), o.stride(0), dout.data_ptr(), dout.stride(0)
    a.lse2, a.D, a.dq_acc, a.ld_dq = lse2.data_ptr(), D.data_ptr(), dq_acc.data_ptr(), C_
    a.dk, a.ld_dk, a.dv, a.ld_dv = dk.data_ptr(), C_, dv.data_ptr(), C_
    a.B, a.heads, a.Nq, a.Nk, a.d, a.scale = B, heads, Nq, Nk, d, float(scale)
    L.check(L.load().unib200_attention_backward(None, C.byref(a), _stream()), "attention_backward")
    dq = torch.empty(B * Nq, C_, device=dev, dtype=torch.float16)
    cvt_f32_f16(dq_acc, dq)
    return dq, dk, dv


def adamw_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, betas, eps: float,
               weight_decay: float, step: int, grad_scale: float = 1.0):
    """torch.optim.AdamW's update on flat fp32 buffers, gradients multiplied by grad_scale first."""
    assert p.dtype == g.dtype == m.dtype == v.dtype == torch.float32 and p.numel() == g.numel() == m.numel() == v.numel()
    L.check(L.load().unib200_adamw_step(None, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, betas[0],
                                        betas[1], eps, weight_decay, step, grad_scale, _stream()), "adamw_step")


class _Linear:
    """y = x W^T (+ b) on the implicit-GEMM kernel, with the two gradient paths of the training slice."""

    def __init__(self, w: torch.Tensor, b: Optional[torch.Tensor], dev):
        self.N, self.K = w.shape
        self.w = ops.pack_weight([(w, SEG_1x1)]).to(dev)
        self.w_t = dgrad_weight(w).to(dev)
        self.b = b.detach().to(dev, torch.float32).contiguous() if b is not None else None

    def fwd(self, x: torch.Tensor, res: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = torch.empty(x.shape[0], self.N, device=x.device, dtype=torch.float16)
        ops.conv_gemm(None, [(x, self.K, SEG_1x1)], self.w, y, M=x.shape[0], N=self.N, bias=self.b, res=res)
        return y

    def bwd(self, x: torch.Tensor, dy: torch.Tensor, partial: torch.Tensor, need_dx: bool = True):
        """(dx | None, dW fp32 [N, K], db fp32 [N] | None)"""
        dw, db = conv_wgrad(x, self.K, dy, self.N, B=1, H=0, W=0, taps=1, partial=partial, want_bias=self.b is not None)
        dx = None
        if need_dx:
            dx = torch.empty(x.shape[0], self.K, device=x.device, dtype=torch.float16)
            ops.conv_gemm(None, [(dy, self.N, SEG_1x1)], self.w_t, dx, M=x.shape[0], N=self.K)
        return dx, dw.reshape(self.N, self.K), db


class TransformerBlockTrainer:
    """BasicTransformerBlock (norm1 -> self-attention -> + ; norm2 -> cross-attention on the text context -> + ;
    norm3 -> GEGLU feed-forward -> +; models/unet_2d_blocks.py:1199-1207 via Transformer2DModel) forward + backward.

    The attention here is the MATERIALISED form (per sample and head: S = Q K^T, row softmax, O = P V through the GEMM
    kernel; P is kept for the backward) -- the fused flash kernel of the inference path has no backward yet, so this is
    a correctness-first path whose score matrices cost Nq x Nk fp16 per head.  Context length must be a multiple of 8."""

    def __init__(self, sd: Dict[str, torch.Tensor], heads: int, device="cuda"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the training slice runs on CUDA (sm_100a) only; there is no CPU fallback")
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()       # noqa: E731
        self.dev, self.heads = dev, heads
        self.C = sd["norm1.weight"].shape[0]
        self.ln = [(f32(sd[f"norm{i}.weight"]), f32(sd[f"norm{i}.bias"])) for i in (1, 2, 3)]
        self.qkv = _Linear(torch.cat([sd["attn1.to_q.weight"], sd["attn1.to_k.weight"], sd["attn1.to_v.weight"]], 0), None, dev)
        self.out1 = _Linear(sd["attn1.to_out.0.weight"], sd["attn1.to_out.0.bias"], dev)
        self.q2 = _Linear(sd["attn2.to_q.weight"], None, dev)
        self.kv2 = _Linear(torch.cat([sd["attn2.to_k.weight"], sd["attn2.to_v.weight"]], 0), None, dev)
        self.out2 = _Linear(sd["attn2.to_out.0.weight"], sd["attn2.to_out.0.bias"], dev)
        self.ffp = _Linear(sd["ff.net.0.proj.weight"], sd["ff.net.0.proj.bias"], dev)
        self.ff2 = _Linear(sd["ff.net.2.weight"], sd["ff.net.2.bias"], dev)
        self.partial = torch.empty(16 << 20, device=dev, dtype=torch.float32)
        self.saved = None

    # -- helpers -------------------------------------------------------------------------------------------------
    @staticmethod
    def _pack_rows(m: torch.Tensor) -> torch.Tensor:
        """[rows, d] column slice -> contiguous [rows, ceil(d/64)*64] (zero padded): the [N, K] operand of a GEMM."""
        rows, d = m.shape
        dst = torch.empty(rows, (d + 63) // 64 * 64, device=m.device, dtype=torch.float16)
        ops.to_nhwc(None, torch.as_strided(m, (1, d, rows, 1), (0, 1, m.stride(0), 1)), dst, dst.shape[1])
        return dst

    @staticmethod
    def _pack_cols(m: torch.Tensor) -> torch.Tensor:
        """[rows, d] column slice -> its transpose, contiguous [d, ceil(rows/64)*64] (zero padded)."""
        rows, d = m.shape
        dst = torch.empty(d, (rows + 63) // 64 * 64, device=m.device, dtype=torch.float16)
        ops.to_nhwc(None, torch.as_strided(m, (1, rows, d, 1), (0, m.stride(0), 1, 1)), dst, dst.shape[1])
        return dst

    def _ln(self, x, i):
        y = torch.empty_like(x)
        ops.layernorm(None, x, y, *self.ln[i])
        return y

    def _attn_fwd(self, q, k, v, B, Nq, Nk):
        """q [B*Nq, C], k / v [B*Nk, C] (column-sliced views allowed) -> (ao [B*Nq, C], P [B, heads, Nq, Nk])."""
        d = self.C // self.heads
        ao = torch.empty(B * Nq, self.C, device=self.dev, dtype=torch.float16)
        P = torch.empty(B, self.heads, Nq, Nk, device=self.dev, dtype=torch.float16)
        for b in range(B):
            for h in range(self.heads):
                cs = slice(h * d, (h + 1) * d)
                qs, ks, vs = q[b * Nq:(b + 1) * Nq, cs], k[b * Nk:(b + 1) * Nk, cs], v[b * Nk:(b + 1) * Nk, cs]
                s = P[b, h]
                ops.conv_gemm(None, [(qs, d, SEG_1x1)], self._pack_rows(ks), s, M=Nq, N=Nk)             # S = Q K^T
                ops.softmax_rows(None, s, rows=Nq, n=Nk, scale=d ** -0.5)
                ops.conv_gemm(None, [(s, Nk, SEG_1x1)], self._pack_cols(vs), ao[b * Nq:(b + 1) * Nq, cs], M=Nq, N=d)
        return ao, P

    def _attn_bwd(self, dao, q, k, v, P, B, Nq, Nk, dq, dk, dv):
        """Gradients of softmax(Q K^T d^-1/2) V into the (column-sliced) matrices dq [B*Nq, C], dk / dv [B*Nk, C]."""
        d = self.C // self.heads
        lib = L.load()
        for b in range(B):
            for h in range(self.heads):
                cs = slice(h * d, (h + 1) * d)
                rq, rk = slice(b * Nq, (b + 1) * Nq), slice(b * Nk, (b + 1) * Nk)
                do, p = dao[rq, cs], P[b, h]
                dvh, _ = conv_wgrad(do, d, p, Nk, B=1, H=0, W=0, taps=1, want_bias=False)               # dV = P^T dO
                L.check(lib.unib200_cvt_f32_f16(None, dvh.data_ptr(), dv[rk, cs].data_ptr(), Nk, d, dv.stride(0), _stream()), "cvt")
                dp = torch.empty(Nq, Nk, device=self.dev, dtype=torch.float16)
                ops.conv_gemm(None, [(do, d, SEG_1x1)], self._pack_rows(v[rk, cs]), dp, M=Nq, N=Nk)    # dP = dO V^T
                L.check(lib.unib200_softmax_backward(None, p.data_ptr(), dp.data_ptr(), Nq, Nk, Nk, d ** -0.5, _stream()),
                        "softmax_backward")                                                            # dp is now dS
                ops.conv_gemm(None, [(dp, Nk, SEG_1x1)], self._pack_cols(k[rk, cs]), dq[rq, cs], M=Nq, N=d)   # dQ = dS K
                dkh, _ = conv_wgrad(q[rq, cs], d, dp, Nk, B=1, H=0, W=0, taps=1, want_bias=False)       # dK = dS^T Q
                L.check(lib.unib200_cvt_f32_f16(None, dkh.data_ptr(), dk[rk, cs].data_ptr(), Nk, d, dk.stride(0), _stream()), "cvt")

    @staticmethod
    def _add(a, b):
        o = torch.empty_like(a)
        ops.add_f16(None, a, b, o)
        return o

    # -- forward / backward ----------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, ctx: torch.Tensor, B: int) -> torch.Tensor:
        """x: fp16 [B*N, C] tokens; ctx: fp16 [B*Lc, Dctx] text context (Lc a multiple of 8)."""
        Cn, N, Lc = self.C, x.shape[0] // B, ctx.shape[0] // B
        n1 = self._ln(x, 0)
        qkv = self.qkv.fwd(n1)
        ao1, P1 = self._attn_fwd(qkv[:, :Cn], qkv[:, Cn:2 * Cn], qkv[:, 2 * Cn:], B, N, N)
        h2 = self.out1.fwd(ao1, res=x)
        n2 = self._ln(h2, 1)
        q2 = self.q2.fwd(n2)
        kv2 = self.kv2.fwd(ctx)
        ao2, P2 = self._attn_fwd(q2, kv2[:, :Cn], kv2[:, Cn:], B, N, Lc)
        h3 = self.out2.fwd(ao2, res=h2)
        n3 = self._ln(h3, 2)
        proj = self.ffp.fwd(n3)
        ffh = geglu(proj)
        out = self.ff2.fwd(ffh, res=h3)
        self.saved = (x, ctx, B, N, Lc, n1, qkv, ao1, P1, h2, n2, q2, kv2, ao2, P2, h3, n3, proj, ffh)
        return out

    def backward(self, dout: torch.Tensor) -> Dict[str, torch.Tensor]:
        x, ctx, B, N, Lc, n1, qkv, ao1, P1, h2, n2, q2, kv2, ao2, P2, h3, n3, proj, ffh = self.saved
        Cn, pt = self.C, self.partial
        g: Dict[str, torch.Tensor] = {}
        # feed-forward
        dffh, g["ff.net.2.weight"], g["ff.net.2.bias"] = self.ff2.bwd(ffh, dout, pt)
        dproj = geglu(proj, dffh)
        dn3, g["ff.net.0.proj.weight"], g["ff.net.0.proj.bias"] = self.ffp.bwd(n3, dproj, pt)
        d3, g["norm3.weight"], g["norm3.bias"] = layernorm_backward(h3, dn3, self.ln[2][0])
        dh3 = self._add(dout, d3)
        # cross-attention
        dao2, g["attn2.to_out.0.weight"], g["attn2.to_out.0.bias"] = self.out2.bwd(ao2, dh3, pt)
        dq2 = torch.empty_like(q2)
        dkv2 = torch.empty_like(kv2)
        self._attn_bwd(dao2, q2, kv2[:, :Cn], kv2[:, Cn:], P2, B, N, Lc, dq2, dkv2[:, :Cn], dkv2[:, Cn:])
        dn2, g["attn2.to_q.weight"], _ = self.q2.bwd(n2, dq2, pt)
        _, dwkv, _ = self.kv2.bwd(ctx, dkv2, pt, need_dx=False)           # the text context is an input without gradient
        g["attn2.to_k.weight"], g["attn2.to_v.weight"] = dwkv[:Cn], dwkv[Cn:]
        d2, g["norm2.weight"], g["norm2.bias"] = layernorm_backward(h2, dn2, self.ln[1][0])
        dh2 = self._add(dh3, d2)
        # self-attention
        dao1, g["attn1.to_out.0.weight"], g["attn1.to_out.0.bias"] = self.out1.bwd(ao1, dh2, pt)
        dqkv = torch.empty_like(qkv)
        self._attn_bwd(dao1, qkv[:, :Cn], qkv[:, Cn:2 * Cn], qkv[:, 2 * Cn:], P1, B, N, N, dqkv[:, :Cn], dqkv[:, Cn:2 * Cn],
                       dqkv[:, 2 * Cn:])
        dn1, dwqkv, _ = self.qkv.bwd(n1, dqkv, pt)
        g["attn1.to_q.weight"], g["attn1.to_k.weight"], g["attn1.to_v.weight"] = dwqkv[:Cn], dwqkv[Cn:2 * Cn], dwqkv[2 * Cn:]
        d1, g["norm1.weight"], g["norm1.bias"] = layernorm_backward(x, dn1, self.ln[0][0])
        g["x"] = self._add(dh2, d1)
        return g
