"""CPU oracle for the AutoencoderKL that brackets the sampling loops.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (uni_renderer_b200/).  Only tests/ may import it.

The reference calls the VAE through the diffusers pipeline object it inherits (paths relative to /root/reference):
    models/pipeline.py:1531-1556   latents_x = self.vae.encode(_x_image).latent_dist.sample() * scaling_factor   (x6-7)
    models/pipeline.py:1664        image = self.vae.decode(latents_img / scaling_factor, return_dict=False)[0]
    models/pipeline.py:2113-2117   (inverse rendering: image + masks encode)   :2335-2344 (4-5 decodes)
The class itself (AutoencoderKL, SD-1.x config: block_out_channels (128, 256, 512, 512), layers_per_block 2,
latent_channels 4, norm_num_groups 32, scaling_factor 0.18215) lives in the third-party dependency
diffusers==0.24.0.dev0 (environment_sam.yml:80), which is neither vendored in the reference nor installable here.
This file restates its published algorithm in plain fp32 PyTorch on a diffusers-layout state dict.

PARITY UNPINNED: unlike oracle/uni_oracle.py there is no reference-side code for this module that could be executed
under oracle/refshim, and the reference holds no golden vectors for it -- the restatement is checked only against
its own structural properties (shapes, key layout, parameter count 83 653 863 of the SD-1.x VAE).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


@dataclass
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    norm_eps: float = 1e-6            # every GroupNorm of the VAE (resnet_eps / attention eps / conv_norm_out)


SD15_VAE = VaeConfig()
TINY_VAE = VaeConfig(block_out_channels=(32, 64, 64), norm_num_groups=8)


def _conv(sd: SD, p: str, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _gn(sd: SD, p: str, x, cfg: VaeConfig):
    return F.group_norm(x, cfg.norm_num_groups, sd[p + ".weight"], sd[p + ".bias"], cfg.norm_eps)


def resnet_block(sd: SD, p: str, x, cfg: VaeConfig):
    """ResnetBlock2D(temb_channels=None, eps=1e-6, output_scale_factor=1): GN-SiLU-conv3x3 twice + (1x1) shortcut."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, cfg)), padding=1)
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, cfg)), padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x)
    return x + h


def attention_block(sd: SD, p: str, x, cfg: VaeConfig):
    """Attention(C, heads=1, dim_head=C, norm_num_groups, eps=1e-6, residual_connection=True, bias=True) as the mid
    block uses it: GroupNorm over the image, q/k/v/out linears WITH bias, one head, softmax(q k^T / sqrt(C)) v."""
    b, c, hh, ww = x.shape
    h = _gn(sd, p + ".group_norm", x, cfg).reshape(b, c, hh * ww).transpose(1, 2)
    q = F.linear(h, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(h, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(h, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    w = torch.softmax((q @ k.transpose(1, 2)) * (c ** -0.5), dim=-1)
    o = F.linear(w @ v, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, hh, ww)


def mid_block(sd: SD, p: str, x, cfg: VaeConfig):
    """UNetMidBlock2D(add_attention=True, attention_head_dim=C): resnet, attention, resnet."""
    x = resnet_block(sd, p + ".resnets.0", x, cfg)
    x = attention_block(sd, p + ".attentions.0", x, cfg)
    return resnet_block(sd, p + ".resnets.1", x, cfg)


def encode_moments(sd: SD, cfg: VaeConfig, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode up to the posterior parameters: Encoder (conv_in, DownEncoderBlock2D x n with
    Downsample2D(padding=0) = F.pad(0,1,0,1) + conv3x3 stride 2, mid block, GN-SiLU-conv_out) + quant_conv.
    Returns moments [B, 2*latent, h, w] = (mean | logvar)."""
    boc = cfg.block_out_channels
    h = _conv(sd, "encoder.conv_in", x, padding=1)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, cfg)
        if i != len(boc) - 1:
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2)
    h = mid_block(sd, "encoder.mid_block", h, cfg)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, cfg)), padding=1)
    return _conv(sd, "quant_conv", h)


def sample_posterior(moments: torch.Tensor, noise: Optional[torch.Tensor]) -> torch.Tensor:
    """DiagonalGaussianDistribution.sample() (noise given) / .mode() (noise None)."""
    mean, logvar = moments.chunk(2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


def decode(sd: SD, cfg: VaeConfig, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode: post_quant_conv + Decoder (conv_in, mid block, UpDecoderBlock2D x n with
    layers_per_block + 1 resnets and nearest-2x + conv3x3 upsamplers, GN-SiLU-conv_out)."""
    boc = cfg.block_out_channels
    rev = list(reversed(boc))
    h = _conv(sd, "post_quant_conv", z)
    h = _conv(sd, "decoder.conv_in", h, padding=1)
    h = mid_block(sd, "decoder.mid_block", h, cfg)
    for i in range(len(rev)):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, cfg)
        if i != len(rev) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h, padding=1)
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, cfg)), padding=1)


# ----------------------------------------------------------------------------------------------------------------
# random-init state dict in the diffusers key layout
# ----------------------------------------------------------------------------------------------------------------
def param_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    sh: Dict[str, Tuple[int, ...]] = {}
    boc = cfg.block_out_channels
    lc = cfg.latent_channels

    def conv(p, i, o, k):
        sh[p + ".weight"] = (o, i, k, k)
        sh[p + ".bias"] = (o,)

    def norm(p, c):
        sh[p + ".weight"] = (c,)
        sh[p + ".bias"] = (c,)

    def lin(p, i, o):
        sh[p + ".weight"] = (o, i)
        sh[p + ".bias"] = (o,)

    def resnet(p, i, o):
        norm(p + ".norm1", i); conv(p + ".conv1", i, o, 3); norm(p + ".norm2", o); conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".conv_shortcut", i, o, 1)

    def mid(p, c):
        resnet(p + ".resnets.0", c, c)
        a = p + ".attentions.0"
        norm(a + ".group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{a}.{n}", c, c)
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
    norm("encoder.conv_norm_out", boc[-1]); conv("encoder.conv_out", boc[-1], 2 * lc, 3)
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
    norm("decoder.conv_norm_out", boc[0]); conv("decoder.conv_out", boc[0], cfg.out_channels, 3)
    return sh


def random_state_dict(cfg: VaeConfig, seed: int) -> SD:
    """torch-default style init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)); norm scales 1 +- 0.1)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    shapes = param_shapes(cfg)
    for name, shp in shapes.items():
        if len(shp) == 1 and "norm" in name:
            t = 0.1 * torch.randn(shp, generator=g)
            if name.endswith("weight"):
                t = t + 1.0
        else:
            wshape = shapes[name[:-len("bias")] + "weight"] if name.endswith("bias") else shp
            k = 1.0 / math.sqrt(math.prod(wshape[1:]))
            t = (torch.rand(shp, generator=g) * 2 - 1) * k
        sd[name] = t
    return sd
