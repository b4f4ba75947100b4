"""CPU oracle for the Uni-Renderer dual-stream denoising hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product path (uni_renderer_b200/).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Plain fp32 PyTorch restatement of the three SD-1.x-shaped networks and the DDIM update, operating directly on
diffusers-layout state dicts (the reference's weight format).  Each function cites the reference lines it follows
(paths relative to /root/reference).  The leaf arithmetic (ResnetBlock2D, Transformer2DModel, Attention, GEGLU,
Down/Upsample2D, Timesteps, TimestepEmbedding, DDIMScheduler) lives in the third-party dependency
diffusers==0.24.0.dev0 (environment_sam.yml:80), which is NOT vendored in the reference and not installable here;
it is restated from the published algorithm.

Parity pinning: the reference holds no golden vectors/tests for this path, so the oracle is
pinned against outputs of the reference's own models/controlnet.py + models/unet_2d_blocks.py executed unmodified
in the build container under oracle/refshim (oracle/make_golden.py -> tests/golden/*.pt; checked by
tests/test_oracle_golden.py).  The diffusers leaves themselves remain "parity unpinned" against real diffusers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


@dataclass
class NetConfig:
    """Subset of the diffusers config the SD-1.x wiring uses (models/controlnet.py:146-205)."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    num_heads: int = 8            # `attention_head_dim` in the reference config (naming quirk, controlnet.py:216-222)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)   # CrossAttnDownBlock2D x3, DownBlock2D
    up_has_attn: Tuple[bool, ...] = (False, True, True, True)     # UpBlock2D, CrossAttnUpBlock2D x3

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


SD15 = NetConfig()
TINY = NetConfig(block_out_channels=(32, 64, 128, 128), num_heads=4, cross_attention_dim=48, norm_num_groups=8)


# ----------------------------------------------------------------------------------------------------------------
# leaves (diffusers 0.24 semantics, restated)
# ----------------------------------------------------------------------------------------------------------------
def timestep_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, freq_shift=0) -- models/controlnet.py:282,909."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    ang = t.reshape(-1, 1).float() * freqs[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _conv(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _gn(sd: SD, p: str, x: torch.Tensor, groups: int, eps: float) -> torch.Tensor:
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps)


def _ln(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def time_embedding(sd: SD, cfg: NetConfig, timestep, batch: Optional[int]) -> torch.Tensor:
    """time_proj + time_embedding.  UNet expands t to the batch (controlnet.py:893-916); the attribute
    encoder/decoder do not (controlnet.py:1682-1708, 2383-2407) so a scalar t gives a (1, D) embedding that
    broadcasts."""
    t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep])
    t = t.reshape(-1)
    if batch is not None:
        t = t.expand(batch)
    e = timestep_sinusoid(t, cfg.block_out_channels[0])
    e = _lin(sd, "time_embedding.linear_1", e)
    e = F.silu(e)
    return _lin(sd, "time_embedding.linear_2", e)


def resnet_block(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, cfg: NetConfig) -> torch.Tensor:
    """ResnetBlock2D(time_embedding_norm="default", output_scale_factor=1) -- constructed at
    models/unet_2d_blocks.py:1100,2458,2619; called at :1199,2575,2696."""
    h = F.silu(_gn(sd, p + ".norm1", x, cfg.norm_num_groups, cfg.norm_eps))
    h = _conv(sd, p + ".conv1", h, padding=1)
    h = h + _lin(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.silu(_gn(sd, p + ".norm2", h, cfg.norm_num_groups, cfg.norm_eps))
    h = _conv(sd, p + ".conv2", h, padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = _conv(sd, p + ".conv_shortcut", x)
    return x + h


def attention(sd: SD, p: str, x: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention (AttnProcessor2_0 -> SDPA): to_q/k/v without bias, to_out.0 with bias, scale d^-1/2."""
    b, n, c = x.shape
    q, k, v = _lin(sd, p + ".to_q", x), _lin(sd, p + ".to_k", ctx), _lin(sd, p + ".to_v", ctx)
    d = c // heads
    q = q.view(b, n, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    w = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    o = (w @ v).transpose(1, 2).reshape(b, n, c)
    return _lin(sd, p + ".to_out.0", o)


def basic_transformer_block(sd: SD, t: str, h: torch.Tensor, ehs: torch.Tensor, heads: int) -> torch.Tensor:
    """BasicTransformerBlock on tokens [B, N, C]: LayerNorm -> self-attention -> +, LayerNorm -> cross-attention on the
    text context -> +, LayerNorm -> GEGLU feed-forward -> + (the body of Transformer2DModel's single layer)."""
    y = _ln(sd, t + ".norm1", h)
    h = attention(sd, t + ".attn1", y, y, heads) + h
    y = _ln(sd, t + ".norm2", h)
    h = attention(sd, t + ".attn2", y, ehs, heads) + h
    y = _ln(sd, t + ".norm3", h)
    a, g = _lin(sd, t + ".ff.net.0.proj", y).chunk(2, dim=-1)
    return _lin(sd, t + ".ff.net.2", a * F.gelu(g)) + h


def transformer_2d(sd: SD, p: str, x: torch.Tensor, ehs: torch.Tensor, cfg: NetConfig) -> torch.Tensor:
    """Transformer2DModel(num_layers=1, use_linear_projection=False) + BasicTransformerBlock + GEGLU FF --
    constructed at models/unet_2d_blocks.py:721,1115,2473; called at :803,1207,2576."""
    b, c, hh, ww = x.shape
    res = x
    h = _gn(sd, p + ".norm", x, cfg.norm_num_groups, 1e-6)
    h = _conv(sd, p + ".proj_in", h)
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    h = basic_transformer_block(sd, p + ".transformer_blocks.0", h, ehs, cfg.num_heads)
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    return _conv(sd, p + ".proj_out", h) + res


# ----------------------------------------------------------------------------------------------------------------
# blocks (models/unet_2d_blocks.py)
# ----------------------------------------------------------------------------------------------------------------
def down_blocks(sd: SD, cfg: NetConfig, sample: torch.Tensor, temb: torch.Tensor, ehs: torch.Tensor):
    """conv_in output + CrossAttnDownBlock2D.forward (:1155-1221) / DownBlock2D.forward (:1276-1309) chain;
    returns (sample, skips) with skips ordered as models/controlnet.py:1051-1073."""
    skips = [sample]
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            sample = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", sample, temb, cfg)
            if cfg.down_has_attn[i]:
                sample = transformer_2d(sd, f"down_blocks.{i}.attentions.{j}", sample, ehs, cfg)
            skips.append(sample)
        if i != nb - 1:
            sample = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", sample, stride=2, padding=1)
            skips.append(sample)
    return sample, skips


def mid_block(sd: SD, cfg: NetConfig, sample, temb, ehs):
    """UNetMidBlock2DCrossAttn.forward (models/unet_2d_blocks.py:764-813)."""
    sample = resnet_block(sd, "mid_block.resnets.0", sample, temb, cfg)
    sample = transformer_2d(sd, "mid_block.attentions.0", sample, ehs, cfg)
    return resnet_block(sd, "mid_block.resnets.1", sample, temb, cfg)


def up_blocks(sd: SD, cfg: NetConfig, sample, skips: Sequence[torch.Tensor], temb, ehs, up_additional=None):
    """CrossAttnUpBlock2D.forward (:2508-2590) / UpBlock2D.forward (:2643-2704) chain; returns (sample, taps)
    where taps mirrors the reference's modified per-resnet outputs (:2584,2697).  `up_additional` (12 tensors, decoder
    layer order) selects the UpResBlock2D / CrossAttnUpResBlock2D variants (models/unet_2d_blocks.py:2706-2822,
    2237-2415): `hidden_states += up_additional_states` after every layer (:2814, :2408)."""
    skips = list(skips)
    up_additional = list(up_additional) if up_additional is not None else None
    taps = [sample]
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            sample = torch.cat([sample, skips.pop()], dim=1)
            sample = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", sample, temb, cfg)
            if cfg.up_has_attn[i]:
                sample = transformer_2d(sd, f"up_blocks.{i}.attentions.{j}", sample, ehs, cfg)
            if up_additional is not None:
                sample = sample + up_additional.pop(0)
            taps.append(sample)
        if i != nb - 1:
            sample = F.interpolate(sample, scale_factor=2.0, mode="nearest")
            sample = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", sample, padding=1)
    return sample, taps


def out_head(sd: SD, cfg: NetConfig, sample):
    """conv_norm_out -> SiLU -> conv_out (models/controlnet.py:1154-1157, 2516-2521)."""
    sample = F.silu(_gn(sd, "conv_norm_out", sample, cfg.norm_num_groups, cfg.norm_eps))
    return _conv(sd, "conv_out", sample, padding=1)


# ----------------------------------------------------------------------------------------------------------------
# the three models (models/controlnet.py)
# ----------------------------------------------------------------------------------------------------------------
def unet_forward(sd: SD, cfg: NetConfig, sample, timestep, ehs, down_block_additional_residuals=None,
                 mid_block_additional_residual=None):
    """UNet2DConditionModel.forward(return_dict=False) -- models/controlnet.py:781-1166.
    Returns (sample, raw_down[12], raw_mid, up_taps[13])."""
    temb = time_embedding(sd, cfg, timestep, sample.shape[0])
    h = _conv(sd, "conv_in", sample, padding=1)
    h, skips = down_blocks(sd, cfg, h, temb, ehs)
    raw_down = tuple(skips)
    is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
    if is_controlnet:   # :1078-1087
        skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
    h = mid_block(sd, cfg, h, temb, ehs)
    raw_mid = h
    if is_controlnet:   # :1114-1115
        h = h + mid_block_additional_residual
    h, taps = up_blocks(sd, cfg, h, skips, temb, ehs)
    return out_head(sd, cfg, h), raw_down, raw_mid, tuple(taps)


def attr_encoder_forward(sd: SD, cfg: NetConfig, timestep, ehs, controlnet_cond, conditioning_scale: float = 1.0):
    """AttributeEncoderModel.forward -- models/controlnet.py:1657-1778.  `sample` is ignored by the reference
    (:1716-1720).  Returns (zero-conv'd down list[12], zero-conv'd mid, raw_down tuple[12], raw_mid)."""
    temb = time_embedding(sd, cfg, timestep, None)
    h = _conv(sd, "conv_in", controlnet_cond, padding=1)
    h, skips = down_blocks(sd, cfg, h, temb, ehs)
    h = mid_block(sd, cfg, h, temb, ehs)
    down = [_conv(sd, f"controlnet_down_blocks.{i}", s) * conditioning_scale for i, s in enumerate(skips)]
    mid = _conv(sd, "controlnet_mid_block", h) * conditioning_scale
    return down, mid, tuple(skips), h


def attr_decoder_forward(sd: SD, cfg: NetConfig, sample, down_block_res_samples, timestep, ehs,
                         down_block_additional_residuals=None, mid_block_additional_residual=None,
                         up_block_additional_residuals=None):
    """AttributeDecoderModel.forward(return_dict=False) -- models/controlnet.py:2342-2527.  With
    `up_block_additional_residuals` the up blocks are the UpRes variants (the class defaults, :1794-1797; the live forward
    has their extra argument commented out at :2486-2510, so this restates the BLOCK semantics of SURVEY row a8)."""
    temb = time_embedding(sd, cfg, timestep, None)
    skips = list(down_block_res_samples)
    if down_block_additional_residuals is not None:    # :2446-2461
        skips = [s + _conv(sd, f"control_down_blocks.{i}", r)
                 for i, (s, r) in enumerate(zip(skips, down_block_additional_residuals))]
    sample = sample + _conv(sd, "control_mid_block", mid_block_additional_residual)   # :2476-2477
    h, _ = up_blocks(sd, cfg, sample, skips, temb, ehs, up_block_additional_residuals)
    return out_head(sd, cfg, h)


def dual_stream_step(sd_unet: SD, sd_enc: SD, sd_dec: SD, cfg_unet: NetConfig, cfg_enc: NetConfig,
                     cfg_dec: NetConfig, x_img, t_img, x_attr, t_attr, ehs):
    """The 3-call dual-stream step: train/train.py:1324-1354 == models/pipeline.py:2660-2690."""
    d, m, raw_a, raw_a_mid = attr_encoder_forward(sd_enc, cfg_enc, t_attr, ehs, x_attr)
    img_pred, raw_u, raw_u_mid, _ = unet_forward(sd_unet, cfg_unet, x_img, t_img, ehs, d, m)
    attr_pred = attr_decoder_forward(sd_dec, cfg_dec, raw_a_mid, raw_a, t_attr, ehs, raw_u, raw_u_mid)
    return img_pred, attr_pred


# ----------------------------------------------------------------------------------------------------------------
# DDIM (diffusers DDIMScheduler with the SD-1.x scheduler config; eta = 0)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class DDIM:
    """DDIMScheduler(beta_start=.00085, beta_end=.012, "scaled_linear", steps_offset=1, set_alpha_to_one=False,
    clip_sample=False, timestep_spacing="leading"); scheduler.step call sites models/pipeline.py:1649,2725-2730,
    models/pipeline_new_d4p.py:1447-1448."""
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    steps_offset: int = 1
    prediction_type: str = "epsilon"
    timestep_spacing: str = "leading"
    beta_schedule: str = "scaled_linear"
    alphas_cumprod: torch.Tensor = field(init=False)
    timesteps: List[int] = field(init=False, default_factory=list)
    num_inference_steps: int = 0

    def __post_init__(self):
        if self.beta_schedule == "scaled_linear":
            betas = torch.linspace(self.beta_start ** 0.5, self.beta_end ** 0.5, self.num_train_timesteps,
                                   dtype=torch.float32) ** 2
        else:
            betas = torch.linspace(self.beta_start, self.beta_end, self.num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def set_timesteps(self, n: int):
        """DDIMScheduler.set_timesteps of diffusers 0.24 for the three `timestep_spacing` values."""
        import numpy as np
        self.num_inference_steps = n
        T = self.num_train_timesteps
        if self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)).astype(np.int64) - 1
        else:
            ts = np.linspace(0, T - 1, n).round()[::-1].copy().astype(np.int64)
        self.timesteps = [int(t) for t in ts]
        return self.timesteps

    def coefficients(self, t: int) -> Tuple[float, float]:
        """x_prev = c_out * model_output + c_x * x_t  (eta=0; every prediction_type is affine in (out, x_t))."""
        a_t = float(self.alphas_cumprod[t])
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.alphas_cumprod[0])
        sa, s1a = math.sqrt(a_t), math.sqrt(1 - a_t)
        sp, s1p = math.sqrt(a_p), math.sqrt(1 - a_p)
        if self.prediction_type == "epsilon":      # x0 = (x - s1a e)/sa ; prev = sp x0 + s1p e
            return s1p - sp * s1a / sa, sp / sa
        if self.prediction_type == "sample":       # e = (x - sa x0)/s1a ; prev = sp x0 + s1p e
            return sp - s1p * sa / s1a, s1p / s1a
        if self.prediction_type == "v_prediction":  # x0 = sa x - s1a v ; e = sa v + s1a x
            return -sp * s1a + s1p * sa, sp * sa + s1p * s1a
        raise ValueError(self.prediction_type)

    def step(self, model_output: torch.Tensor, t: int, sample: torch.Tensor) -> torch.Tensor:
        a_t = self.alphas_cumprod[t]
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.alphas_cumprod[0]
        if self.prediction_type == "epsilon":
            x0 = (sample - (1 - a_t).sqrt() * model_output) / a_t.sqrt()
            eps = model_output
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t.sqrt() * x0) / (1 - a_t).sqrt()
        elif self.prediction_type == "v_prediction":
            x0 = a_t.sqrt() * sample - (1 - a_t).sqrt() * model_output
            eps = a_t.sqrt() * model_output + (1 - a_t).sqrt() * sample
        else:
            raise ValueError(self.prediction_type)
        return a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps


@dataclass
class UniPC:
    """UniPCMultistepScheduler as the shipped eval drives it (eval/test_real.py:485-493: one scheduler per stream,
    `UniPCMultistepScheduler.from_config(pipeline.scheduler.config)`, 20 steps; step call sites
    models/pipeline.py:1649,2725-2730).  Restated from the published UniPC algorithm (Zhao et al. 2023, "UniPC: A
    Unified Predictor-Corrector Framework") in the form diffusers 0.24 implements it: solver_order 2, solver_type
    "bh2", predict_x0, lower_order_final, no thresholding, sigmas interpolated at the (linspace-spaced) timesteps with
    the final sigma appended.  PARITY UNPINNED against diffusers itself (not installable here) -- it pins the
    product's closed-form coefficient tables (uni_renderer_b200/scheduler.py) through an independent formulation:
    this class runs the tensor algorithm step by step with its model-output history."""
    num_train_timesteps: int = 1000
    beta_start: float = 0.00085
    beta_end: float = 0.012
    prediction_type: str = "epsilon"
    solver_order: int = 2
    timestep_spacing: str = "linspace"
    steps_offset: int = 0
    solver_type: str = "bh2"
    beta_schedule: str = "scaled_linear"

    def __post_init__(self):
        if self.beta_schedule == "scaled_linear":
            betas = torch.linspace(self.beta_start ** 0.5, self.beta_end ** 0.5, self.num_train_timesteps,
                                   dtype=torch.float64) ** 2
        else:
            betas = torch.linspace(self.beta_start, self.beta_end, self.num_train_timesteps, dtype=torch.float64)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def set_timesteps(self, n: int):
        import numpy as np
        T = self.num_train_timesteps
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":       # what from_config(SD-1.x scheduler config) inherits
            ts = (np.arange(0, n + 1) * (T // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        else:
            ts = np.arange(T, 0, -T / n).round().copy().astype(np.int64) - 1
        sig = ((1 - self.alphas_cumprod) / self.alphas_cumprod).sqrt().numpy()
        s = np.interp(ts, np.arange(T), sig)
        s_last = float(((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]).sqrt())
        self.sigmas = [float(v) for v in s] + [s_last]
        self.timesteps = [int(t) for t in ts]
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.step_index = 0
        self.last_sample = None
        self.this_order = 1
        return self.timesteps

    @staticmethod
    def _alpha_sigma(sigma: float):
        alpha = 1.0 / math.sqrt(sigma * sigma + 1.0)
        return alpha, sigma * alpha

    def _convert(self, model_output, sample):
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[self.step_index])
        if self.prediction_type == "epsilon":
            return (sample - sigma_t * model_output) / alpha_t
        if self.prediction_type == "sample":
            return model_output
        if self.prediction_type == "v_prediction":
            return alpha_t * sample - sigma_t * model_output
        raise ValueError(self.prediction_type)

    def _bh(self, s_t: float, s_s0: float, s_hist, order: int):
        """Shared scalar part of the predictor / corrector: returns (alpha_t, sigma_t/sigma_s0, h_phi_1, B_h, rks, R, b)."""
        alpha_t, sigma_t = self._alpha_sigma(s_t)
        alpha_s0, sigma_s0 = self._alpha_sigma(s_s0)
        lam_t, lam_s0 = math.log(alpha_t) - math.log(sigma_t), math.log(alpha_s0) - math.log(sigma_s0)
        h = lam_t - lam_s0
        rks = []
        for s_i in s_hist:
            a_i, sg_i = self._alpha_sigma(s_i)
            rks.append(((math.log(a_i) - math.log(sg_i)) - lam_s0) / h)
        rks.append(1.0)
        hh = -h                                   # predict_x0
        h_phi_1 = math.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1.0
        B_h = math.expm1(hh) if self.solver_type == "bh2" else hh
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append([rk ** (i - 1) for rk in rks])
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1.0 / fact
        return alpha_t, sigma_t / sigma_s0, h_phi_1, B_h, rks, R, b

    def _predict(self, sample, order: int):
        m0 = self.model_outputs[-1]
        hist = [self.sigmas[self.step_index - i] for i in range(1, order)]
        alpha_t, ratio, h_phi_1, B_h, rks, R, b = self._bh(self.sigmas[self.step_index + 1], self.sigmas[self.step_index],
                                                           hist, order)
        D1s = [(self.model_outputs[-(i + 1)] - m0) / rks[i - 1] for i in range(1, order)]
        x_t_ = ratio * sample - alpha_t * h_phi_1 * m0
        if D1s:
            if order == 2:
                rhos_p = [0.5]
            else:
                rhos_p = torch.linalg.solve(torch.tensor(R, dtype=torch.float64)[:-1, :-1],
                                            torch.tensor(b, dtype=torch.float64)[:-1]).tolist()
            pred = sum(r * d for r, d in zip(rhos_p, D1s))
        else:
            pred = 0.0
        return x_t_ - alpha_t * B_h * pred

    def _correct(self, model_t, last_sample, order: int):
        m0 = self.model_outputs[-1]
        hist = [self.sigmas[self.step_index - (i + 1)] for i in range(1, order)]
        alpha_t, ratio, h_phi_1, B_h, rks, R, b = self._bh(self.sigmas[self.step_index], self.sigmas[self.step_index - 1],
                                                           hist, order)
        D1s = [(self.model_outputs[-(i + 1)] - m0) / rks[i - 1] for i in range(1, order)]
        if order == 1:
            rhos_c = [0.5]
        else:
            rhos_c = torch.linalg.solve(torch.tensor(R, dtype=torch.float64),
                                        torch.tensor(b, dtype=torch.float64)).tolist()
        x_t_ = ratio * last_sample - alpha_t * h_phi_1 * m0
        corr = sum(r * d for r, d in zip(rhos_c[:-1], D1s)) if D1s else 0.0
        return x_t_ - alpha_t * B_h * (corr + rhos_c[-1] * (model_t - m0))

    def step(self, model_output: torch.Tensor, t: int, sample: torch.Tensor) -> torch.Tensor:
        assert t == self.timesteps[self.step_index], "steps must be taken in schedule order"
        use_corrector = self.step_index > 0 and self.last_sample is not None
        x0 = self._convert(model_output, sample)
        if use_corrector:
            sample = self._correct(x0, self.last_sample, self.this_order)
        self.model_outputs = self.model_outputs[1:] + [x0]
        this_order = min(self.solver_order, len(self.timesteps) - self.step_index)         # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


# ----------------------------------------------------------------------------------------------------------------
# random-init state dicts in the reference's key layout (for synthetic benchmarks / tests without the reference)
# ----------------------------------------------------------------------------------------------------------------
def _init_linear(sd: SD, p: str, cin: int, cout: int, g: torch.Generator, bias: bool = True):
    k = 1.0 / math.sqrt(cin)
    sd[p + ".weight"] = (torch.rand(cout, cin, generator=g) * 2 - 1) * k
    if bias:
        sd[p + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * k


def _init_conv(sd: SD, p: str, cin: int, cout: int, ks: int, g: torch.Generator):
    k = 1.0 / math.sqrt(cin * ks * ks)
    sd[p + ".weight"] = (torch.rand(cout, cin, ks, ks, generator=g) * 2 - 1) * k
    sd[p + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * k


def _init_norm(sd: SD, p: str, c: int, g: torch.Generator):
    sd[p + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
    sd[p + ".bias"] = 0.1 * torch.randn(c, generator=g)


def _init_resnet(sd: SD, p: str, cin: int, cout: int, cfg: NetConfig, g):
    _init_norm(sd, p + ".norm1", cin, g)
    _init_conv(sd, p + ".conv1", cin, cout, 3, g)
    _init_linear(sd, p + ".time_emb_proj", cfg.time_embed_dim, cout, g)
    _init_norm(sd, p + ".norm2", cout, g)
    _init_conv(sd, p + ".conv2", cout, cout, 3, g)
    if cin != cout:
        _init_conv(sd, p + ".conv_shortcut", cin, cout, 1, g)


def _init_transformer(sd: SD, p: str, c: int, cfg: NetConfig, g):
    _init_norm(sd, p + ".norm", c, g)
    _init_conv(sd, p + ".proj_in", c, c, 1, g)
    t = p + ".transformer_blocks.0"
    for n in ("norm1", "norm2", "norm3"):
        _init_norm(sd, f"{t}.{n}", c, g)
    for a, kv in (("attn1", c), ("attn2", cfg.cross_attention_dim)):
        _init_linear(sd, f"{t}.{a}.to_q", c, c, g, bias=False)
        _init_linear(sd, f"{t}.{a}.to_k", kv, c, g, bias=False)
        _init_linear(sd, f"{t}.{a}.to_v", kv, c, g, bias=False)
        _init_linear(sd, f"{t}.{a}.to_out.0", c, c, g)
    _init_linear(sd, f"{t}.ff.net.0.proj", c, 8 * c, g)
    _init_linear(sd, f"{t}.ff.net.2", 4 * c, c, g)
    _init_conv(sd, p + ".proj_out", c, c, 1, g)


def skip_channels(cfg: NetConfig) -> List[int]:
    """Channel width of the 12 skips in reference order (models/controlnet.py:1051-1073)."""
    ch = [cfg.block_out_channels[0]]
    nb = len(cfg.block_out_channels)
    for i, c in enumerate(cfg.block_out_channels):
        ch += [c] * cfg.layers_per_block
        if i != nb - 1:
            ch.append(c)
    return ch


def random_state_dict(kind: str, cfg: NetConfig, seed: int, zero_conv_std: float = 0.02) -> SD:
    """kind in {"unet", "attr_enc", "attr_dec"}; key layout per SURVEY.md section 8b.  The zero-convs are drawn
    N(0, zero_conv_std) instead of zeros so the exchange is numerically live."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    boc = cfg.block_out_channels
    nb = len(boc)
    _init_linear(sd, "time_embedding.linear_1", boc[0], cfg.time_embed_dim, g)
    _init_linear(sd, "time_embedding.linear_2", cfg.time_embed_dim, cfg.time_embed_dim, g)
    if kind in ("unet", "attr_enc"):
        _init_conv(sd, "conv_in", cfg.in_channels, boc[0], 3, g)
        cin = boc[0]
        for i, c in enumerate(boc):
            for j in range(cfg.layers_per_block):
                _init_resnet(sd, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c, cfg, g)
                if cfg.down_has_attn[i]:
                    _init_transformer(sd, f"down_blocks.{i}.attentions.{j}", c, cfg, g)
            if i != nb - 1:
                _init_conv(sd, f"down_blocks.{i}.downsamplers.0.conv", c, c, 3, g)
            cin = c
        _init_resnet(sd, "mid_block.resnets.0", boc[-1], boc[-1], cfg, g)
        _init_transformer(sd, "mid_block.attentions.0", boc[-1], cfg, g)
        _init_resnet(sd, "mid_block.resnets.1", boc[-1], boc[-1], cfg, g)
    if kind in ("unet", "attr_dec"):
        rev = list(reversed(boc))
        prev = rev[0]
        for i, c in enumerate(rev):
            cin_skip = rev[min(i + 1, nb - 1)]
            for j in range(cfg.layers_per_block + 1):
                skip = cin_skip if j == cfg.layers_per_block else c
                rin = prev if j == 0 else c
                _init_resnet(sd, f"up_blocks.{i}.resnets.{j}", rin + skip, c, cfg, g)
                if cfg.up_has_attn[i]:
                    _init_transformer(sd, f"up_blocks.{i}.attentions.{j}", c, cfg, g)
            if i != nb - 1:
                _init_conv(sd, f"up_blocks.{i}.upsamplers.0.conv", c, c, 3, g)
            prev = c
        _init_norm(sd, "conv_norm_out", boc[0], g)
        _init_conv(sd, "conv_out", boc[0], cfg.out_channels, 3, g)
    if kind in ("attr_enc", "attr_dec"):
        name = "controlnet" if kind == "attr_enc" else "control"
        for i, c in enumerate(skip_channels(cfg)):
            sd[f"{name}_down_blocks.{i}.weight"] = torch.randn(c, c, 1, 1, generator=g) * zero_conv_std
            sd[f"{name}_down_blocks.{i}.bias"] = torch.randn(c, generator=g) * zero_conv_std
        sd[f"{name}_mid_block.weight"] = torch.randn(boc[-1], boc[-1], 1, 1, generator=g) * zero_conv_std
        sd[f"{name}_mid_block.bias"] = torch.randn(boc[-1], generator=g) * zero_conv_std
    return sd
