"""AutoencoderKL around the sampling loops, on the B200 kernels.

The reference brackets every sampling call with the diffusers VAE of the pipeline object (models/pipeline.py):
    :1531-1556  latents_x = self.vae.encode(_x_image).latent_dist.sample() * self.vae.config.scaling_factor   (x6-7)
    :1664       image = self.vae.decode(latents_img / self.vae.config.scaling_factor, return_dict=False)[0]
    :2113-2117  inverse rendering: encode(image), encode(masks);  :2335-2344  four to five decodes of the attributes
At 512x512 the five decodes of inverse rendering cost about as much as the 20-step loop once the UNets are fast.

`AutoencoderKL` here keeps the diffusers surface those call sites use (`encode(x).latent_dist.sample()/.mode()`,
`decode(z, return_dict=False)[0]`, `.config.scaling_factor`, `.config.block_out_channels`, `.dtype`,
`enable/disable_slicing/tiling`, `from_pretrained(dir, subfolder="vae")`, the diffusers state-dict keys) and runs
every layer as a recorded program of the same hand-written kernels as the UNets: implicit-GEMM tcgen05 convolutions
(3x3, the bottom/right-padded stride-2 downsample as SEG_3x3_S2P0, 1x1 shortcuts accumulated into conv2's TMEM
accumulator), GroupNorm+SiLU, nearest-2x upsample.  quant_conv is folded into the encoder's conv_out weights (exact:
a 1x1 conv after a 3x3 conv is a 3x3 conv).  The single-head d = C attention of the mid block does not fit the flash
kernel's TMEM layout; it runs as three GEMMs around an in-place row softmax (include/unib200.h).

No CPU / eager fallback: without CUDA or libunib200.so every call raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch
from torch import nn

from . import ops
from .engine import Act, Workspace, _f32, pad_channels
from .models import _Config, _NetModule
from . import _lib as L
from .ops import EPI_OUT_F32, EPI_OUT_NCHW, SEG_1x1, SEG_3x3, SEG_3x3_S2P0, SEG_UP2x2

_DOWN = "DownEncoderBlock2D"
_UP = "UpDecoderBlock2D"


@dataclass
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    norm_eps: float = 1e-6


def vae_param_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of every AutoencoderKL parameter in the diffusers (>= 0.20) key layout."""
    sh: Dict[str, Tuple[int, ...]] = {}
    boc, lc = cfg.block_out_channels, cfg.latent_channels

    def conv(p, i, o, k):
        sh[p + ".weight"] = (o, i, k, k)
        sh[p + ".bias"] = (o,)

    def vec(p, c):
        sh[p + ".weight"] = (c,)
        sh[p + ".bias"] = (c,)

    def resnet(p, i, o):
        vec(p + ".norm1", i); conv(p + ".conv1", i, o, 3); vec(p + ".norm2", o); conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".conv_shortcut", i, o, 1)

    def mid(p, c):
        resnet(p + ".resnets.0", c, c)
        vec(p + ".attentions.0.group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            sh[f"{p}.attentions.0.{n}.weight"] = (c, c)
            sh[f"{p}.attentions.0.{n}.bias"] = (c,)
        resnet(p + ".resnets.1", c, c)

    conv("encoder.conv_in", cfg.in_channels, boc[0], 3)
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c)
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
        cin = c
    mid("encoder.mid_block", boc[-1])
    vec("encoder.conv_norm_out", boc[-1]); conv("encoder.conv_out", boc[-1], 2 * lc, 3)
    conv("quant_conv", 2 * lc, 2 * lc, 1)
    conv("post_quant_conv", lc, lc, 1)
    rev = list(reversed(boc))
    conv("decoder.conv_in", lc, rev[0], 3)
    mid("decoder.mid_block", rev[0])
    cin = rev[0]
    for i, c in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else c, c)
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
        cin = c
    vec("decoder.conv_norm_out", boc[0]); conv("decoder.conv_out", boc[0], cfg.out_channels, 3)
    return sh


# pre-0.20 checkpoints name the mid-block attention parameters like the original LDM code; diffusers renames them at
# load time (Attention._from_deprecated_attn_block), and so do we
_DEPRECATED_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


def convert_deprecated_attention_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if len(parts) >= 3 and parts[-3] == "0" and "attentions" in parts and parts[-2] in _DEPRECATED_ATTN:
            parts[-2:-1] = _DEPRECATED_ATTN[parts[-2]].split(".")
            if v.dim() == 4:                       # very old checkpoints store the projections as 1x1 convs
                v = v[:, :, 0, 0]
        out[".".join(parts)] = v
    return out


def fold_quant_conv(w_out: torch.Tensor, b_out: torch.Tensor, w_q: torch.Tensor, b_q: torch.Tensor):
    """quant_conv(conv_out(x)): a 1x1 conv applied to the result of a 3x3 conv is the 3x3 conv with weights
    W'[o, i, ky, kx] = sum_m Wq[o, m] Wc[m, i, ky, kx] and bias Wq bc + bq (exact; zero padding commutes because the
    1x1 acts on the OUTPUT of the padded conv)."""
    wq = w_q.detach().float().reshape(w_q.shape[0], w_q.shape[1])
    w = torch.einsum("om,mikl->oikl", wq, w_out.detach().float())
    b = wq @ b_out.detach().float() + b_q.detach().float()
    return w, b


class VaeNet:
    """Packed weights of one AutoencoderKL + the recorders of its two halves."""

    def __init__(self, cfg: VaeConfig, sd: Dict[str, torch.Tensor], device):
        self.cfg, self.device = cfg, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("uni_renderer_b200 runs on CUDA (sm_100a) only -- there is no CPU fallback")
        L.load()
        c_mid = cfg.block_out_channels[-1]
        if c_mid % 64:
            raise ValueError("the mid-block attention needs block_out_channels[-1] to be a multiple of 64")
        if any(c % 8 or c % cfg.norm_num_groups for c in cfg.block_out_channels):
            raise ValueError("block_out_channels must be multiples of 8 and of norm_num_groups")
        self.w: Dict[str, torch.Tensor] = {}
        self._pack(convert_deprecated_attention_keys(sd))

    def _pack(self, sd):
        import os
        from .engine import gn_granularity
        # GroupNorm statistics from the producing GEMMs' epilogues + folded upsampling, as in the UNets (engine.py)
        self.gn_gran = gn_granularity(self.cfg.block_out_channels, self.cfg.norm_num_groups)
        self.gn_min_hw, self.gn_force, self._sms = 256, os.environ.get("UNIB200_GN_FUSED") == "force", None
        self.upfold = os.environ.get("UNIB200_UPFOLD", "1") != "0"
        dev, w = self.device, self.w
        need = lambda k: sd[k].detach().to(dev)    # noqa: E731

        def conv3(name, kind=SEG_3x3):
            w[name + ".w"] = ops.pack_weight([(need(name + ".weight"), kind)])
            w[name + ".b"] = _f32(sd[name + ".bias"], dev)

        def norm(name):
            w[name + ".g"] = _f32(sd[name + ".weight"], dev)
            w[name + ".bt"] = _f32(sd[name + ".bias"], dev)

        for r in [k[:-len(".conv1.weight")] for k in sd if k.endswith(".conv1.weight")]:
            norm(r + ".norm1"); norm(r + ".norm2"); conv3(r + ".conv1")
            parts, bias = [(need(r + ".conv2.weight"), SEG_3x3)], sd[r + ".conv2.bias"].detach().to(dev).float()
            has_sc = r + ".conv_shortcut.weight" in sd
            if has_sc:                              # 1x1 shortcut accumulated into conv2's accumulator
                parts.append((need(r + ".conv_shortcut.weight"), SEG_1x1))
                bias = bias + sd[r + ".conv_shortcut.bias"].detach().to(dev).float()
            w[r + ".conv2.w"] = ops.pack_weight(parts)
            w[r + ".conv2.b"] = bias.contiguous()
            w[r + ".has_sc"] = has_sc
        for m in ("encoder.mid_block.attentions.0", "decoder.mid_block.attentions.0"):
            norm(m + ".group_norm")
            for n in ("to_q", "to_k", "to_out.0"):
                w[f"{m}.{n}.w"] = ops.pack_weight([(need(f"{m}.{n}.weight"), SEG_1x1)])
                w[f"{m}.{n}.b"] = _f32(sd[f"{m}.{n}.bias"], dev)
            # to_v is used as the A operand of the V^T GEMM (rows = output channels): plain fp16 [C, C]
            w[m + ".to_v.a"] = need(m + ".to_v.weight").half().contiguous()
            w[m + ".to_v.b"] = _f32(sd[m + ".to_v.bias"], dev)
        nb = len(self.cfg.block_out_channels)
        conv3("encoder.conv_in")
        for i in range(nb - 1):
            conv3(f"encoder.down_blocks.{i}.downsamplers.0.conv", SEG_3x3_S2P0)
            n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            if self.upfold and ops.upfold_supported(sd[n + ".weight"].shape[0]):
                w[n + ".wup"] = ops.pack_upsample_conv(need(n + ".weight"))      # nearest-2x folded into the conv
                w[n + ".b"] = _f32(sd[n + ".bias"], dev)
            else:
                conv3(n)
        norm("encoder.conv_norm_out")
        wq, bq = fold_quant_conv(need("encoder.conv_out.weight"), need("encoder.conv_out.bias"),
                                 need("quant_conv.weight"), need("quant_conv.bias"))
        w["encoder.conv_out.w"] = ops.pack_weight([(wq, SEG_3x3)])
        w["encoder.conv_out.b"] = bq.contiguous()
        w["post_quant_conv.w"] = ops.pack_weight([(need("post_quant_conv.weight"), SEG_1x1)])
        w["post_quant_conv.b"] = _f32(sd["post_quant_conv.bias"], dev)
        conv3("decoder.conv_in")
        norm("decoder.conv_norm_out")
        conv3("decoder.conv_out")

    # ------------------------------------------------------------------------------------------------------------
    def gn_plan(self, M: int, N: int, B: int, HW: int):
        from .engine import plan_gn_stats
        return plan_gn_stats(self, M, N, B, HW)

    def _gn(self, prog, ws, name, x: Act, silu: bool) -> torch.Tensor:
        out = ws.get(x.M, x.C)
        parts = None
        if x.gn is not None and (x.C // self.cfg.norm_num_groups) % x.gn[1] == 0:
            parts = (x.gn[0], None, x.gn[1], x.gn[2])
        ops.groupnorm(prog, x.t, x.C, None, 0, self.w[name + ".g"], self.w[name + ".bt"], out, ws.gn_scratch, B=x.B,
                      HW=x.H * x.W, groups=self.cfg.norm_num_groups, eps=self.cfg.norm_eps, silu=silu, parts=parts)
        return out

    def rec_resnet(self, prog, ws, r: str, x: Act) -> Act:
        """ResnetBlock2D(temb_channels=None): GN-SiLU-conv1, GN-SiLU-conv2 (+ shortcut / identity residual)."""
        B, H, W, M = x.B, x.H, x.W, x.M
        Cout = self.w[r + ".conv1.w"].shape[0]
        n1 = self._gn(prog, ws, r + ".norm1", x, True)
        h1 = ws.get(M, Cout)
        gn1 = self.gn_plan(M, Cout, B, H * W)
        ops.conv_gemm(prog, [(n1, x.C, SEG_3x3)], self.w[r + ".conv1.w"], h1, M=M, N=Cout, B=B, H=H, W=W,
                      bias=self.w[r + ".conv1.b"], partial=None if gn1 else ws.partial, gn=gn1)
        ws.put(n1)
        n2 = self._gn(prog, ws, r + ".norm2", Act(h1, B, H, W, Cout, gn1), True)
        ws.put(h1)
        out = ws.get(M, Cout)
        segs, res = [(n2, Cout, SEG_3x3)], None
        if self.w[r + ".has_sc"]:
            segs.append((x.t, x.C, SEG_1x1))
        else:
            res = x.t
        gn2 = self.gn_plan(M, Cout, B, H * W)
        ops.conv_gemm(prog, segs, self.w[r + ".conv2.w"], out, M=M, N=Cout, B=B, H=H, W=W, bias=self.w[r + ".conv2.b"],
                      res=res, partial=None if gn2 else ws.partial, gn=gn2)
        ws.put(n2)
        return Act(out, B, H, W, Cout, gn2)

    def rec_attention(self, prog, ws, a: str, x: Act) -> Act:
        """x + to_out(softmax(q k^T / sqrt(C)) v) with q, k, v = linear(GroupNorm(x)): one head of d = C."""
        B, T, Cc, M = x.B, x.H * x.W, x.C, x.M
        if T % 64:
            raise ValueError("the VAE mid-block attention needs H*W to be a multiple of 64 at the latent resolution")
        w = self.w
        g = self._gn(prog, ws, a + ".group_norm", x, False)
        q, k, ao = ws.get(M, Cc), ws.get(M, Cc), ws.get(M, Cc)
        ops.conv_gemm(prog, [(g, Cc, SEG_1x1)], w[a + ".to_q.w"], q, M=M, N=Cc, B=B, bias=w[a + ".to_q.b"])
        ops.conv_gemm(prog, [(g, Cc, SEG_1x1)], w[a + ".to_k.w"], k, M=M, N=Cc, B=B, bias=w[a + ".to_k.b"])
        s = ws.get(T, T)                                   # score / probability matrix of ONE sample, reused
        vt = ws.get(Cc, T)                                 # V^T of one sample
        for b in range(B):
            rows = slice(b * T, (b + 1) * T)
            # V^T[c, n] = sum_k Wv[c, k] g[n, k]: the weight matrix is the A operand, the activation the [N, K] one.
            # The bias of to_v is added after P V instead (rows of P sum to 1).
            ops.conv_gemm(prog, [(w[a + ".to_v.a"], Cc, SEG_1x1)], g[rows], vt, M=Cc, N=T)
            ops.conv_gemm(prog, [(q[rows], Cc, SEG_1x1)], k[rows], s, M=T, N=T)              # S = Q K^T
            ops.softmax_rows(prog, s, rows=T, n=T, scale=Cc ** -0.5)
            ops.conv_gemm(prog, [(s, T, SEG_1x1)], vt, ao[rows], M=T, N=Cc, bias=w[a + ".to_v.b"])   # O = P V + bv
        out = ws.get(M, Cc)
        ops.conv_gemm(prog, [(ao, Cc, SEG_1x1)], w[a + ".to_out.0.w"], out, M=M, N=Cc, B=B, bias=w[a + ".to_out.0.b"],
                      res=x.t)
        ws.put(g, q, k, ao, s, vt)
        return Act(out, B, x.H, x.W, Cc)

    def rec_mid(self, prog, ws, p: str, x: Act) -> Act:
        r0 = self.rec_resnet(prog, ws, p + ".resnets.0", x)
        at = self.rec_attention(prog, ws, p + ".attentions.0", r0)
        ws.put(r0.t)
        r1 = self.rec_resnet(prog, ws, p + ".resnets.1", at)
        ws.put(at.t)
        return r1

    def rec_encoder(self, prog, ws, x_in: Act, moments_nchw: torch.Tensor):
        """image (NHWC fp16, channels padded to 8) -> posterior moments fp32 NCHW [B, 2*latent, h, w]."""
        cfg = self.cfg
        boc = cfg.block_out_channels
        B = x_in.B
        h = Act(ws.get(x_in.M, boc[0]), B, x_in.H, x_in.W, boc[0])
        h.gn = self.gn_plan(h.M, boc[0], B, h.H * h.W)
        ops.conv_gemm(prog, [(x_in.t, x_in.C, SEG_3x3)], self.w["encoder.conv_in.w"], h.t, M=h.M, N=boc[0], B=B,
                      H=h.H, W=h.W, bias=self.w["encoder.conv_in.b"], gn=h.gn)
        for i in range(len(boc)):
            for j in range(cfg.layers_per_block):
                r = self.rec_resnet(prog, ws, f"encoder.down_blocks.{i}.resnets.{j}", h)
                ws.put(h.t)
                h = r
            if i != len(boc) - 1:
                n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                o = Act(ws.get(h.M // 4, h.C), B, h.H // 2, h.W // 2, h.C)
                o.gn = self.gn_plan(o.M, o.C, B, o.H * o.W)
                ops.conv_gemm(prog, [(h.t, h.C, SEG_3x3_S2P0)], self.w[n + ".w"], o.t, M=o.M, N=o.C, B=B, H=o.H, W=o.W,
                              bias=self.w[n + ".b"], partial=None if o.gn else ws.partial, gn=o.gn)
                ws.put(h.t)
                h = o
        m = self.rec_mid(prog, ws, "encoder.mid_block", h)
        ws.put(h.t)
        g = self._gn(prog, ws, "encoder.conv_norm_out", m, True)
        ops.conv_gemm(prog, [(g, m.C, SEG_3x3)], self.w["encoder.conv_out.w"], moments_nchw, M=m.M,
                      N=2 * cfg.latent_channels, B=B, H=m.H, W=m.W, bias=self.w["encoder.conv_out.b"],
                      flags=EPI_OUT_NCHW | EPI_OUT_F32)
        ws.put(g, m.t)

    def rec_decoder(self, prog, ws, z_in: Act, image_nchw: torch.Tensor):
        """latents (NHWC fp16, channels padded to 8; already divided by scaling_factor) -> image fp32 NCHW."""
        cfg = self.cfg
        rev = list(reversed(cfg.block_out_channels))
        B, lc = z_in.B, cfg.latent_channels
        # post_quant_conv: a 4 -> 4 1x1 conv.  It cannot be folded into conv_in (its bias would leak into conv_in's
        # zero padding), so it runs as its own tiny GEMM through the NCHW fp32 epilogue and is re-laid out.
        pq = torch.empty(B, lc, z_in.H, z_in.W, device=self.device, dtype=torch.float32)
        ops.conv_gemm(prog, [(z_in.t, z_in.C, SEG_1x1)], self.w["post_quant_conv.w"], pq, M=z_in.M, N=lc, B=B,
                      bias=self.w["post_quant_conv.b"], flags=EPI_OUT_NCHW | EPI_OUT_F32)
        zq = Act(torch.zeros(z_in.M, z_in.C, device=self.device, dtype=torch.float16), B, z_in.H, z_in.W, z_in.C)
        ops.to_nhwc(prog, pq, zq.t, zq.C)
        h = Act(ws.get(zq.M, rev[0]), B, zq.H, zq.W, rev[0])
        h.gn = self.gn_plan(h.M, rev[0], B, h.H * h.W)
        ops.conv_gemm(prog, [(zq.t, zq.C, SEG_3x3)], self.w["decoder.conv_in.w"], h.t, M=h.M, N=rev[0], B=B, H=h.H,
                      W=h.W, bias=self.w["decoder.conv_in.b"], partial=None if h.gn else ws.partial, gn=h.gn)
        m = self.rec_mid(prog, ws, "decoder.mid_block", h)
        ws.put(h.t)
        h = m
        for i in range(len(rev)):
            for j in range(cfg.layers_per_block + 1):
                r = self.rec_resnet(prog, ws, f"decoder.up_blocks.{i}.resnets.{j}", h)
                ws.put(h.t)
                h = r
            if i != len(rev) - 1:
                n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                o = Act(ws.get(h.M * 4, h.C), B, h.H * 2, h.W * 2, h.C)
                if n + ".wup" in self.w:            # Upsample2D as ONE GEMM over the low-resolution tensor (SEG_UP2x2)
                    gnu = self.gn_plan(o.M, o.C, B, o.H * o.W)
                    if gnu is not None and (gnu[2] != 128 or (h.H * h.W) % 128):
                        gnu = None
                    o.gn = gnu
                    ops.conv_gemm(prog, [(h.t, h.C, SEG_UP2x2)], self.w[n + ".wup"], o.t, M=h.M, N=4 * o.C, B=B, H=h.H,
                                  W=h.W, bias=self.w[n + ".b"], gn=gnu)
                    ws.put(h.t)
                else:
                    up = ws.get(h.M * 4, h.C)
                    ops.upsample2x(prog, h.t, up, B=B, H=h.H, W=h.W, Cn=h.C)
                    o.gn = self.gn_plan(o.M, o.C, B, o.H * o.W)
                    ops.conv_gemm(prog, [(up, h.C, SEG_3x3)], self.w[n + ".w"], o.t, M=o.M, N=o.C, B=B, H=o.H, W=o.W,
                                  bias=self.w[n + ".b"], partial=None if o.gn else ws.partial, gn=o.gn)
                    ws.put(up, h.t)
                h = o
        g = self._gn(prog, ws, "decoder.conv_norm_out", h, True)
        ops.conv_gemm(prog, [(g, h.C, SEG_3x3)], self.w["decoder.conv_out.w"], image_nchw, M=h.M, N=cfg.out_channels,
                      B=B, H=h.H, W=h.W, bias=self.w["decoder.conv_out.b"], flags=EPI_OUT_NCHW | EPI_OUT_F32)
        ws.put(g, h.t)


# ----------------------------------------------------------------------------------------------------------------
# drop-in module surface
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class AutoencoderKLOutput:
    latent_dist: "DiagonalGaussianDistribution" = None


@dataclass
class DecoderOutput:
    sample: torch.Tensor = None


class DiagonalGaussianDistribution:
    """`vae.encode(x).latent_dist`: holds the fp32 moments [B, 2C, h, w]; sample() / mode() run on the device."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = parameters.chunk(2, dim=1)

    def _draw(self, noise: Optional[torch.Tensor]) -> torch.Tensor:
        B, C2, H, W = self.parameters.shape
        out = torch.empty(B, C2 // 2, H, W, device=self.parameters.device, dtype=torch.float32)
        ops.gaussian_sample(None, self.parameters, noise, out, scale=1.0)
        return out

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        B, C2, H, W = self.parameters.shape
        dev = self.parameters.device
        gdev = generator.device if generator is not None else dev
        noise = torch.randn(B, C2 // 2, H, W, generator=generator, device=gdev, dtype=torch.float32).to(dev)
        return self._draw(noise)

    def mode(self) -> torch.Tensor:
        return self._draw(None)


class AutoencoderKL(_NetModule):
    """diffusers.AutoencoderKL surface over VaeNet (inference only)."""
    _kind = "vae"
    max_batch = 8          # images per recorded program; larger batches are processed in chunks of this size
    use_graph = os.environ.get("UNIB200_VAE_GRAPH", "1") != "0"     # replay each program as one CUDA graph

    def __init__(self, in_channels: int = 3, out_channels: int = 3, down_block_types=(_DOWN,),
                 up_block_types=(_UP,), block_out_channels=(64,), layers_per_block: int = 1, act_fn: str = "silu",
                 latent_channels: int = 4, norm_num_groups: int = 32, sample_size: int = 32,
                 scaling_factor: float = 0.18215, force_upcast: bool = True, _init_weights: bool = True, **unused):
        cfg_kwargs = dict(in_channels=in_channels, out_channels=out_channels, down_block_types=tuple(down_block_types),
                          up_block_types=tuple(up_block_types), block_out_channels=tuple(block_out_channels),
                          layers_per_block=layers_per_block, act_fn=act_fn, latent_channels=latent_channels,
                          norm_num_groups=norm_num_groups, sample_size=sample_size, scaling_factor=scaling_factor,
                          force_upcast=force_upcast)
        nn.Module.__init__(self)
        if any(t != _DOWN for t in down_block_types) or any(t != _UP for t in up_block_types):
            raise ValueError("only DownEncoderBlock2D / UpDecoderBlock2D blocks are supported")
        if not (len(down_block_types) == len(up_block_types) == len(block_out_channels)):
            raise ValueError("down_block_types, up_block_types and block_out_channels must have the same length")
        if act_fn not in ("silu", "swish"):
            raise ValueError(f"act_fn={act_fn!r} is not supported")
        self.vae_cfg = VaeConfig(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                                 block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                 norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self._internal_dict = _Config(cfg_kwargs)
        for name, shp in vae_param_shapes(self.vae_cfg).items():
            t = torch.empty(shp, dtype=torch.float32)
            if _init_weights:
                self._init_param(name, t)
            self._add_param(name, t)
        self._net: Optional[VaeNet] = None
        self._progs: Dict[Any, Any] = {}
        self._ws: Optional[Workspace] = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        return super().load_state_dict(convert_deprecated_attention_keys(state_dict), strict=strict, assign=assign)

    def finalize(self, device=None) -> VaeNet:
        if self._net is None:
            device = torch.device(device) if device is not None else self.device
            if device.type != "cuda":
                raise RuntimeError("uni_renderer_b200 runs on CUDA (sm_100a) only; move the module to a CUDA device -- "
                                   "there is no CPU fallback")
            self._net = VaeNet(self.vae_cfg, dict(self.state_dict()), device)
            self._ws = Workspace(device)
        return self._net

    # the reference pipeline toggles these (models/pipeline.py:190-215); batches are processed whole here
    def enable_slicing(self):
        return None

    def disable_slicing(self):
        return None

    def enable_tiling(self, *a, **k):
        raise NotImplementedError("tiled VAE decoding changes the result (blended tiles) and is not implemented")

    def disable_tiling(self):
        return None

    def _program(self, which: str, B: int, H: int, W: int):
        key = (which, B, H, W)
        if key in self._progs:
            return self._progs[key]
        net, ws, cfg, dev = self.finalize(), self._ws, self.vae_cfg, self._net.device
        f = 2 ** (len(cfg.block_out_channels) - 1)
        P = {"prog": ops.Program()}
        if which == "enc":
            if H % f or W % f or (H & (H - 1)) or (W & (W - 1)):
                raise ValueError(f"image sides must be powers of two (got {H}x{W})")
            cp = pad_channels(cfg.in_channels)
            P["x"] = Act(torch.zeros(B * H * W, cp, device=dev, dtype=torch.float16), B, H, W, cp)
            P["moments"] = torch.zeros(B, 2 * cfg.latent_channels, H // f, W // f, device=dev, dtype=torch.float32)
            net.rec_encoder(P["prog"], ws, P["x"], P["moments"])
        else:
            if (H & (H - 1)) or (W & (W - 1)):
                raise ValueError(f"latent sides must be powers of two (got {H}x{W})")
            cp = pad_channels(cfg.latent_channels)
            P["z"] = Act(torch.zeros(B * H * W, cp, device=dev, dtype=torch.float16), B, H, W, cp)
            P["image"] = torch.zeros(B, cfg.out_channels, H * f, W * f, device=dev, dtype=torch.float32)
            net.rec_decoder(P["prog"], ws, P["z"], P["image"])
        if self.use_graph:
            # one eager pass first (sets every kernel's function attributes outside capture), then capture on a side
            # stream -- the legacy default stream cannot be captured (same recipe as DualStreamSampler.plan)
            P["prog"].run()
            torch.cuda.synchronize(dev)
            side = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(side):
                P["prog"].instantiate_graph()
            side.synchronize()
        self._progs[key] = P
        return P

    @staticmethod
    def _replay(P):
        P["prog"].launch_graph() if P["prog"].has_graph else P["prog"].run()

    @staticmethod
    def _float(x: torch.Tensor) -> torch.Tensor:
        return x if x.dtype in (torch.float16, torch.float32) else x.float()

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """`self.vae.encode(image).latent_dist.sample()` (models/pipeline.py:1531, 2113).  The returned moments are a
        fresh tensor (the program's static buffer is copied out), so several encodes can be alive at once."""
        net = self.finalize(x.device if x.is_cuda else None)
        B, _, H, W = x.shape
        if B > self.max_batch:         # large batches run as chunks of one recorded program (bounded activation memory)
            parts = [self.encode(x[i:i + self.max_batch]).latent_dist.parameters for i in range(0, B, self.max_batch)]
            dist = DiagonalGaussianDistribution(torch.cat(parts, 0))
            return (dist,) if not return_dict else AutoencoderKLOutput(latent_dist=dist)
        P = self._program("enc", B, H, W)
        ops.to_nhwc(None, self._float(x).to(net.device), P["x"].t, P["x"].C)
        self._replay(P)
        dist = DiagonalGaussianDistribution(P["moments"].clone())
        if not return_dict:
            return (dist,)
        return AutoencoderKLOutput(latent_dist=dist)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        """`self.vae.decode(latents / scaling_factor, return_dict=False)[0]` (models/pipeline.py:1664, 2335-2344).
        `generator` is accepted and unused, as in diffusers.  Returns a fresh fp32 (or z.dtype) NCHW image."""
        net = self.finalize(z.device if z.is_cuda else None)
        B, _, H, W = z.shape
        if B > self.max_batch:
            img = torch.cat([self.decode(z[i:i + self.max_batch], return_dict=False)[0]
                             for i in range(0, B, self.max_batch)], 0)
            return (img,) if not return_dict else DecoderOutput(sample=img)
        P = self._program("dec", B, H, W)
        ops.to_nhwc(None, self._float(z).to(net.device), P["z"].t, P["z"].C)
        self._replay(P)
        img = P["image"].to(z.dtype) if z.dtype in (torch.float16, torch.bfloat16) else P["image"].clone()
        if not return_dict:
            return (img,)
        return DecoderOutput(sample=img)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None):
        post = self.encode(sample).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z, return_dict=return_dict)


def random_init_vae_state_dict(cfg: VaeConfig, seed: int, device="cuda", dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Random-init AutoencoderKL weights in the diffusers key layout, generated on `device` (synthetic benchmarks)."""
    g = torch.Generator(device=device).manual_seed(seed)
    shapes = vae_param_shapes(cfg)
    sd: Dict[str, torch.Tensor] = {}
    for name, shp in shapes.items():
        if len(shp) == 1 and "norm" in name:
            t = 0.1 * torch.randn(shp, generator=g, device=device, dtype=torch.float32)
            if name.endswith("weight"):
                t += 1.0
        else:
            wshape = shapes[name[:-len("bias")] + "weight"] if name.endswith("bias") else shp
            fan_in = 1
            for s in wshape[1:]:
                fan_in *= s
            t = (torch.rand(shp, generator=g, device=device, dtype=torch.float32) * 2 - 1) * fan_in ** -0.5
        sd[name] = t.to(dtype)
    return sd
