"""Drop-in module surface for the three networks of the dual-stream step.

Same class names, constructor keywords, `forward()` signatures, return structures, `.config` / `.dtype` /
`from_unet()` / state-dict key layout as the reference (models/controlnet.py: UNet2DConditionModel :49,
AttributeEncoderModel :1170, AttributeDecoderModel :1781), so `train/train.py:1324-1354`-style and
`models/pipeline.py:2660-2690`-style callers run unchanged -- but every forward replays a recorded program of
hand-written sm_100a kernels (uni_renderer_b200.engine).  Inference only (no autograd); requires a CUDA device and the
built libunib200.so, otherwise it raises: there is no eager fallback.

Returned tensors are channels_last-strided fp16 VIEWS of the module's static activation buffers (valid until the next
forward() of the same module with the same input shape); the reference returns fresh tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch
from torch import nn

from . import ops
from .engine import Act, NetConfig, StreamNet, Workspace, pad_channels

_SD_DOWN = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D")
_SD_UP = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
# the class defaults of AttributeDecoderModel (models/controlnet.py:1794-1797): same parameters as _SD_UP, but every layer
# adds an extra residual (`hidden_states += up_additional_states`, models/unet_2d_blocks.py:2408,2814 -- SURVEY row a8)
_SD_UPRES = ("UpResBlock2D", "CrossAttnUpResBlock2D", "CrossAttnUpResBlock2D", "CrossAttnUpResBlock2D")


@dataclass
class UNet2DConditionOutput:
    """Same single-field output object as diffusers' (models/controlnet.py:1166)."""
    sample: torch.Tensor = None


class _Config(dict):
    """Mutable attribute dict, like diffusers' FrozenDict-backed `.config` as the reference uses it
    (`unet.config.in_channels`, `"x" in unet.config`, `config["in_channels"] = ...` at train/train.py:985,996)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _ParamTree(nn.Module):
    """Nested parameter containers so state_dict() keys equal the diffusers layout."""

    def add(self, dotted: str, tensor: torch.Tensor):
        head, _, rest = dotted.partition(".")
        if not rest:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=False))
            return
        if head not in self._modules:
            self.add_module(head, _ParamTree())
        self._modules[head].add(rest, tensor)


def _param_shapes(kind: str, cfg: NetConfig) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of every parameter in the reference layout."""
    sh: Dict[str, Tuple[int, ...]] = {}
    boc, D, X = cfg.block_out_channels, cfg.time_embed_dim, cfg.cross_attention_dim
    nb = len(boc)

    def lin(p, i, o, bias=True):
        sh[p + ".weight"] = (o, i)
        if bias:
            sh[p + ".bias"] = (o,)

    def conv(p, i, o, k):
        sh[p + ".weight"] = (o, i, k, k)
        sh[p + ".bias"] = (o,)

    def norm(p, c):
        sh[p + ".weight"] = (c,)
        sh[p + ".bias"] = (c,)

    def resnet(p, i, o):
        norm(p + ".norm1", i); conv(p + ".conv1", i, o, 3); lin(p + ".time_emb_proj", D, o)
        norm(p + ".norm2", o); conv(p + ".conv2", o, o, 3)
        if i != o:
            conv(p + ".conv_shortcut", i, o, 1)

    def transformer(p, c):
        norm(p + ".norm", c); conv(p + ".proj_in", c, c, 1)
        t = p + ".transformer_blocks.0"
        for n in ("norm1", "norm2", "norm3"):
            norm(f"{t}.{n}", c)
        for a, kvd in (("attn1", c), ("attn2", X)):
            lin(f"{t}.{a}.to_q", c, c, False); lin(f"{t}.{a}.to_k", kvd, c, False); lin(f"{t}.{a}.to_v", kvd, c, False)
            lin(f"{t}.{a}.to_out.0", c, c)
        lin(f"{t}.ff.net.0.proj", c, 8 * c); lin(f"{t}.ff.net.2", 4 * c, c)
        conv(p + ".proj_out", c, c, 1)

    lin("time_embedding.linear_1", boc[0], D); lin("time_embedding.linear_2", D, D)
    if kind in ("unet", "attr_enc"):
        conv("conv_in", cfg.in_channels, boc[0], 3)
        cin = boc[0]
        for i, c in enumerate(boc):
            for j in range(cfg.layers_per_block):
                resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else c, c)
                if cfg.down_has_attn[i]:
                    transformer(f"down_blocks.{i}.attentions.{j}", c)
            if i != nb - 1:
                conv(f"down_blocks.{i}.downsamplers.0.conv", c, c, 3)
            cin = c
        resnet("mid_block.resnets.0", boc[-1], boc[-1]); transformer("mid_block.attentions.0", boc[-1])
        resnet("mid_block.resnets.1", boc[-1], boc[-1])
    if kind in ("unet", "attr_dec"):
        rev = list(reversed(boc))
        prev = rev[0]
        for i, c in enumerate(rev):
            cskip = rev[min(i + 1, nb - 1)]
            for j in range(cfg.layers_per_block + 1):
                resnet(f"up_blocks.{i}.resnets.{j}", (prev if j == 0 else c) + (cskip if j == cfg.layers_per_block else c), c)
                if cfg.up_has_attn[i]:
                    transformer(f"up_blocks.{i}.attentions.{j}", c)
            if i != nb - 1:
                conv(f"up_blocks.{i}.upsamplers.0.conv", c, c, 3)
            prev = c
        norm("conv_norm_out", boc[0]); conv("conv_out", boc[0], cfg.out_channels, 3)
    if kind in ("attr_enc", "attr_dec"):
        zc = "controlnet" if kind == "attr_enc" else "control"
        ch = [boc[0]]
        for i, c in enumerate(boc):
            ch += [c] * cfg.layers_per_block + ([c] if i != nb - 1 else [])
        for i, c in enumerate(ch):
            conv(f"{zc}_down_blocks.{i}", c, c, 1)
        conv(f"{zc}_mid_block", boc[-1], boc[-1], 1)
    return sh


def _as_act(t: torch.Tensor) -> Optional[Act]:
    """Zero-copy view of a [B, C, H, W] fp16 CUDA tensor whose memory is NHWC with a 16-byte-aligned pixel stride."""
    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float16 and t.dim() == 4):
        return None
    B, Cn, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    if sc != 1 or sw < Cn or sw % 8 or sh != W * sw or sb != H * W * sw or t.data_ptr() % 16:
        return None
    return Act(torch.as_strided(t, (B * H * W, Cn), (sw, 1)), B, H, W, Cn)


class _NetModule(nn.Module):
    """Shared machinery: config handling, parameter tree, lazy packing, per-shape recorded programs."""
    _kind = "unet"
    _supports_gradient_checkpointing = True

    def _init_common(self, cfg_kwargs: Dict[str, Any], net_cfg: NetConfig, init_weights: bool = True):
        self._internal_dict = _Config(cfg_kwargs)
        self.net_cfg = net_cfg
        shapes = _param_shapes(self._kind, net_cfg)
        for name, shp in shapes.items():
            t = torch.empty(shp, dtype=torch.float32)
            if init_weights:
                self._init_param(name, t)
            self._add_param(name, t)
        self._net: Optional[StreamNet] = None
        self._progs: Dict[Any, Any] = {}
        self._ws: Optional[Workspace] = None

    # parameters live directly on self so state_dict keys have no prefix
    def _add_param(self, dotted: str, tensor: torch.Tensor):
        head, _, rest = dotted.partition(".")
        if not rest:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=False))
            return
        if head not in self._modules:
            self.add_module(head, _ParamTree())
        self._modules[head].add(rest, tensor)

    @staticmethod
    def _init_param(name: str, t: torch.Tensor):
        with torch.no_grad():
            if name.startswith(("controlnet_", "control_")):
                t.zero_()                                           # zero_module(), controlnet.py:1360-1415,1988-2009
            elif t.dim() == 1:
                if name.endswith("weight") and ("norm" in name):
                    t.fill_(1.0)
                else:
                    t.zero_()
            else:
                fan_in = t[0].numel()
                t.uniform_(-fan_in ** -0.5, fan_in ** -0.5)

    @property
    def config(self) -> _Config:
        return self._internal_dict

    @property
    def dtype(self) -> torch.dtype:
        """Compute/storage dtype of the B200 path (weights are packed to fp16, accumulation is fp32)."""
        return torch.float16

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def _invalidate(self):
        self._net, self._progs, self._ws = None, {}, None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._invalidate()
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    # -- diffusers-style persistence (train/train.py:905-916,1470-1476 load and save the three networks this way) ----
    WEIGHTS_NAME = "diffusion_pytorch_model.bin"
    SAFETENSORS_WEIGHTS_NAME = "diffusion_pytorch_model.safetensors"
    CONFIG_NAME = "config.json"

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True, **unused):
        """`<dir>/config.json` + `<dir>/diffusion_pytorch_model.safetensors` (or `.bin`), the layout diffusers'
        ModelMixin.save_pretrained writes and `from_pretrained` reads."""
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        cfg = {"_class_name": type(self).__name__, "_diffusers_version": "0.24.0.dev0"}
        cfg.update({k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()})
        with open(os.path.join(save_directory, self.CONFIG_NAME), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        sd = {k: v.detach().cpu().contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, self.SAFETENSORS_WEIGHTS_NAME), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, self.WEIGHTS_NAME))

    @classmethod
    def load_config(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **unused) -> Dict[str, Any]:
        import json
        import os
        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        path = os.path.join(d, cls.CONFIG_NAME)
        if not os.path.isfile(path):
            raise EnvironmentError(f"{path} not found (only local directories are supported: there is no hub access)")
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config: Dict[str, Any], **overrides):
        """Construct from a diffusers config dict: keys starting with '_' and keys the constructor does not know are
        dropped (diffusers' ConfigMixin.extract_init_dict does the same)."""
        import inspect
        init_weights = overrides.pop("_init_weights", True)
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self", "unused", "_init_weights"}
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in dict(config, **overrides).items()
              if not k.startswith("_") and k in accepted}
        return cls(**kw, _init_weights=init_weights)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None,
                        torch_dtype: Optional[torch.dtype] = None, **unused):
        """Local-directory subset of diffusers' ModelMixin.from_pretrained (train/train.py:905-916:
        `UNet2DConditionModel.from_pretrained(path, subfolder="unet")`): config.json -> constructor, then the
        safetensors / .bin state dict with strict key checking."""
        import os
        cfg = cls.load_config(pretrained_model_name_or_path, subfolder=subfolder)
        m = cls.from_config(cfg, _init_weights=False)
        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        st, pt = os.path.join(d, cls.SAFETENSORS_WEIGHTS_NAME), os.path.join(d, cls.WEIGHTS_NAME)
        if os.path.isfile(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        elif os.path.isfile(pt):
            sd = torch.load(pt, map_location="cpu", weights_only=True)
        else:
            raise EnvironmentError(f"no {cls.SAFETENSORS_WEIGHTS_NAME} or {cls.WEIGHTS_NAME} in {d}")
        m.load_state_dict(sd, strict=True)
        if torch_dtype is not None:
            m = m.to(torch_dtype)
        return m

    def enable_xformers_memory_efficient_attention(self, *a, **k):   # reference callers invoke these; no-ops here
        return None

    # -- the rest of the reference's module surface (models/controlnet.py:629 set_attn_processor, :680
    #    set_attention_slice, :649 set_default_attn_processor, :749 enable_freeu, :773 disable_freeu, :591
    #    attn_processors): the attention and the skip path are fixed fused kernels here, so the switches that would
    #    change them raise a clear error instead of an AttributeError; the ones that restore the default are no-ops.
    @property
    def attn_processors(self) -> Dict[str, Any]:
        return {}

    def set_attn_processor(self, processor, _remove_lora: bool = False):
        raise NotImplementedError("set_attn_processor: attention runs in the fused tcgen05 kernel of uni_renderer_b200; "
                                  "custom attention processors (LoRA, added-KV, sliced) are not supported")

    def set_default_attn_processor(self):
        return None

    def set_attention_slice(self, slice_size):
        raise NotImplementedError("set_attention_slice: the fused attention kernel never materialises the score matrix, "
                                  "slicing is neither needed nor supported")

    def enable_freeu(self, s1, s2, b1, b2):
        raise NotImplementedError("enable_freeu: FreeU rescales the skip / backbone features inside the up blocks "
                                  "(models/unet_2d_blocks.py:2522-2544); no shipped Uni-Renderer caller enables it and "
                                  "the fused decoder does not implement it")

    def disable_freeu(self):
        return None

    def enable_gradient_checkpointing(self):
        raise NotImplementedError("uni_renderer_b200 is an inference path: training/backward is out of scope")

    def finalize(self, device=None) -> StreamNet:
        """Pack the current parameters for the kernels (called lazily by forward)."""
        if self._net is None:
            device = torch.device(device) if device is not None else self.device
            if device.type != "cuda":
                raise RuntimeError("uni_renderer_b200 runs on CUDA (sm_100a) only; move the module to a CUDA device -- "
                                   "there is no CPU fallback")
            sd = {k: v for k, v in self.state_dict().items()}
            self._net = StreamNet(self._kind, self.net_cfg, sd, device)
            self._ws = Workspace(device)
        return self._net

    # -- helpers --------------------------------------------------------------------------------------------------
    @staticmethod
    def _timesteps(timestep, B: int, device) -> torch.Tensor:
        """python number | 0-d | (1,) | (B,) -> fp32 [B] (controlnet.py:893-907; the attribute nets broadcast a
        scalar t, :1682-1708, which is the same thing as expanding it)."""
        if not torch.is_tensor(timestep):
            t = torch.tensor([float(timestep)], dtype=torch.float32)
        else:
            t = timestep.detach().reshape(-1).to(torch.float32)
        if t.numel() == 1:
            t = t.expand(B)
        if t.numel() != B:
            raise ValueError(f"timestep has {t.numel()} entries for batch {B}")
        return t.to(device)

    @staticmethod
    def _reject(**kw):
        bad = [k for k, v in kw.items() if v is not None]
        if bad:
            raise NotImplementedError(f"arguments {bad} are not supported by the B200 hot path "
                                      "(no shipped Uni-Renderer caller passes them)")


def _net_config(in_channels, out_channels, block_out_channels, layers_per_block, attention_head_dim,
                cross_attention_dim, norm_num_groups, norm_eps, down_block_types, up_block_types, **other) -> NetConfig:
    if tuple(down_block_types) != _SD_DOWN[:len(block_out_channels)] and tuple(down_block_types) != _SD_DOWN:
        raise ValueError(f"unsupported down_block_types {down_block_types}")
    if tuple(up_block_types) not in (_SD_UP, _SD_UPRES):
        raise ValueError(f"unsupported up_block_types {up_block_types}")
    if not isinstance(attention_head_dim, int) or not isinstance(layers_per_block, int):
        raise ValueError("per-block attention_head_dim / layers_per_block tuples are not supported")
    if norm_num_groups is None:
        raise ValueError("norm_num_groups=None is not supported")
    unsupported = {"center_input_sample": (False,), "only_cross_attention": (False,), "dual_cross_attention": (False,),
                   "use_linear_projection": (False,), "class_embed_type": (None,), "addition_embed_type": (None,),
                   "num_class_embeds": (None,), "upcast_attention": (False,), "resnet_time_scale_shift": ("default",),
                   "encoder_hid_dim": (None,), "act_fn": ("silu", "swish"), "mid_block_scale_factor": (1, 1.0),
                   "transformer_layers_per_block": (1,), "global_pool_conditions": (False,)}
    for k, v in other.items():
        if k in unsupported and v not in unsupported[k]:
            raise ValueError(f"config option {k}={v!r} is not supported by the B200 hot path "
                             f"(supported: {unsupported[k]})")
    return NetConfig(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                     layers_per_block=layers_per_block, num_heads=attention_head_dim,
                     cross_attention_dim=cross_attention_dim, norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                     up_res=tuple(up_block_types) == _SD_UPRES)


class UNet2DConditionModel(_NetModule):
    """RGB-stream UNet (reference: models/controlnet.py:49; forward :781-1166)."""
    _kind = "unet"

    def __init__(self, sample_size=None, in_channels: int = 8, out_channels: int = 8, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0, down_block_types=_SD_DOWN, mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=_SD_UP, only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1, dropout=0.0, act_fn="silu",
                 norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280, transformer_layers_per_block=1,
                 attention_head_dim=8, num_attention_heads=None, dual_cross_attention=False,
                 use_linear_projection=False, class_embed_type=None, addition_embed_type=None, num_class_embeds=None,
                 upcast_attention=False, resnet_time_scale_shift="default", time_embedding_type="positional",
                 encoder_hid_dim=None, encoder_hid_dim_type=None, _init_weights: bool = True, **unused):
        super().__init__()
        if num_attention_heads is not None:
            raise ValueError("num_attention_heads cannot be set (same restriction as the reference, controlnet.py:212)")
        cfg_kwargs = dict(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                          center_input_sample=center_input_sample, flip_sin_to_cos=flip_sin_to_cos,
                          freq_shift=freq_shift, down_block_types=tuple(down_block_types), mid_block_type=mid_block_type,
                          up_block_types=tuple(up_block_types), only_cross_attention=only_cross_attention,
                          block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                          downsample_padding=downsample_padding, mid_block_scale_factor=mid_block_scale_factor,
                          dropout=dropout, act_fn=act_fn, norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                          cross_attention_dim=cross_attention_dim,
                          transformer_layers_per_block=transformer_layers_per_block,
                          attention_head_dim=attention_head_dim, num_attention_heads=num_attention_heads,
                          dual_cross_attention=dual_cross_attention, use_linear_projection=use_linear_projection,
                          class_embed_type=class_embed_type, addition_embed_type=addition_embed_type,
                          addition_time_embed_dim=None, num_class_embeds=num_class_embeds,
                          upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift,
                          time_embedding_type=time_embedding_type, encoder_hid_dim=encoder_hid_dim,
                          encoder_hid_dim_type=encoder_hid_dim_type, projection_class_embeddings_input_dim=None)
        nc = _net_config(in_channels, out_channels, block_out_channels, layers_per_block, attention_head_dim,
                         cross_attention_dim, norm_num_groups, norm_eps, down_block_types, up_block_types,
                         center_input_sample=center_input_sample, only_cross_attention=only_cross_attention,
                         dual_cross_attention=dual_cross_attention, use_linear_projection=use_linear_projection,
                         class_embed_type=class_embed_type, addition_embed_type=addition_embed_type,
                         num_class_embeds=num_class_embeds, upcast_attention=upcast_attention,
                         resnet_time_scale_shift=resnet_time_scale_shift, encoder_hid_dim=encoder_hid_dim,
                         act_fn=act_fn, mid_block_scale_factor=mid_block_scale_factor,
                         transformer_layers_per_block=transformer_layers_per_block)
        self._init_common(cfg_kwargs, nc, _init_weights)

    def _program(self, B, H, W, L, with_res: bool):
        key = (B, H, W, L, with_res)
        if key in self._progs:
            return self._progs[key]
        net, ws, cfg, dev = self.finalize(), self._ws, self.net_cfg, self._net.device
        P = {"prog": ops.Program()}
        prog = P["prog"]
        P["t"] = torch.zeros(B, device=dev, dtype=torch.float32)
        P["ehs"] = torch.zeros(B * L, cfg.cross_attention_dim, device=dev, dtype=torch.float16)
        cp = pad_channels(cfg.in_channels)
        P["x"] = Act(torch.zeros(B * H * W, cp, device=dev, dtype=torch.float16), B, H, W, cp)
        P["out"] = torch.zeros(B, cfg.out_channels, H, W, device=dev, dtype=torch.float32)
        tproj = net.rec_temb(prog, ws, P["t"], B)
        kv = net.rec_kv(prog, ws, P["ehs"], B, L)
        skips, mid = net.rec_encoder(prog, ws, P["x"], tproj, kv, L)
        P["raw_down"], P["raw_mid"] = skips, mid
        dskips, dmid = skips, mid
        if with_res:
            P["res_in"] = [Act(torch.zeros_like(s.t), s.B, s.H, s.W, s.C) for s in skips]
            P["res_mid_in"] = Act(torch.zeros_like(mid.t), mid.B, mid.H, mid.W, mid.C)
            dskips = []
            for s, r in zip(skips, P["res_in"]):
                o = Act(torch.empty_like(s.t), s.B, s.H, s.W, s.C)
                ops.add_f16(prog, s.t, r.t, o.t)                      # controlnet.py:1078-1087
                dskips.append(o)
            dmid = Act(torch.empty_like(mid.t), mid.B, mid.H, mid.W, mid.C)
            ops.add_f16(prog, mid.t, P["res_mid_in"].t, dmid.t)       # controlnet.py:1114-1115
        taps: list = []
        net.rec_decoder(prog, ws, dmid, dskips, tproj, kv, L, out_nchw=P["out"], taps=taps)
        P["taps"] = taps
        self._progs[key] = P
        return P

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                down_intrablock_additional_residuals=None, encoder_attention_mask=None, return_dict: bool = True):
        self._reject(class_labels=class_labels, timestep_cond=timestep_cond, attention_mask=attention_mask,
                     added_cond_kwargs=added_cond_kwargs, encoder_attention_mask=encoder_attention_mask,
                     down_intrablock_additional_residuals=down_intrablock_additional_residuals)
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs (LoRA scale / GLIGEN) are not supported")
        B, _, H, W = sample.shape
        L = encoder_hidden_states.shape[1]
        with_res = down_block_additional_residuals is not None and mid_block_additional_residual is not None
        if down_block_additional_residuals is not None and not with_res:
            raise NotImplementedError("T2I-adapter style residuals without a mid residual are not supported")
        P = self._program(B, H, W, L, with_res)
        dev = P["t"].device
        P["t"].copy_(self._timesteps(timestep, B, dev))
        P["ehs"].copy_(encoder_hidden_states.reshape(B * L, -1))
        x = sample
        if self.config.get("center_input_sample"):
            x = 2 * x - 1.0
        ops.to_nhwc(None, x if x.dtype in (torch.float16, torch.float32) else x.float(), P["x"].t, P["x"].C)
        if with_res:
            for slot, r in zip(P["res_in"], down_block_additional_residuals):
                slot.nchw().copy_(r)
            P["res_mid_in"].nchw().copy_(mid_block_additional_residual)
        P["prog"].run()
        out = P["out"] if sample.dtype == torch.float32 else P["out"].to(sample.dtype)
        if not return_dict:
            return (out, tuple(s.nchw() for s in P["raw_down"]), P["raw_mid"].nchw(), tuple(t.nchw() for t in P["taps"]))
        return UNet2DConditionOutput(sample=out)


class AttributeEncoderModel(_NetModule):
    """Attribute-stream encoder + 13 zero-convs (reference: models/controlnet.py:1170; forward :1657-1778)."""
    _kind = "attr_enc"

    def __init__(self, in_channels: int = 4, conditioning_channels: int = 3, flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=_SD_DOWN, only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32,
                 norm_eps=1e-5, cross_attention_dim=1280, transformer_layers_per_block=1, encoder_hid_dim=None,
                 encoder_hid_dim_type=None, attention_head_dim=8, num_attention_heads=None, use_linear_projection=False,
                 class_embed_type=None, addition_embed_type=None, addition_time_embed_dim=None, num_class_embeds=None,
                 upcast_attention=False, resnet_time_scale_shift="default",
                 projection_class_embeddings_input_dim=None, controlnet_conditioning_channel_order="rgb",
                 conditioning_embedding_out_channels=(16, 32, 96, 256), global_pool_conditions=False,
                 addition_embed_type_num_heads=64, len_t: int = 1, _init_weights: bool = True):
        cfg_kwargs = dict(locals())
        for k in ("self", "__class__", "_init_weights"):
            cfg_kwargs.pop(k, None)
        super().__init__()
        cfg_kwargs["down_block_types"] = tuple(down_block_types)
        cfg_kwargs["block_out_channels"] = tuple(block_out_channels)
        if num_attention_heads not in (None, attention_head_dim):
            raise ValueError("num_attention_heads must be None")
        nc = _net_config(in_channels, 0, block_out_channels, layers_per_block, attention_head_dim, cross_attention_dim,
                         norm_num_groups, norm_eps, down_block_types, _SD_UP, only_cross_attention=only_cross_attention,
                         use_linear_projection=use_linear_projection, class_embed_type=class_embed_type,
                         addition_embed_type=addition_embed_type, num_class_embeds=num_class_embeds,
                         upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift,
                         encoder_hid_dim=encoder_hid_dim, act_fn=act_fn, mid_block_scale_factor=mid_block_scale_factor,
                         transformer_layers_per_block=transformer_layers_per_block,
                         global_pool_conditions=global_pool_conditions)
        self.len_t = len_t
        self._init_common(cfg_kwargs, nc, _init_weights)

    @classmethod
    def from_unet(cls, unet: UNet2DConditionModel, controlnet_conditioning_channel_order: str = "rgb",
                  conditioning_embedding_out_channels=(16, 32, 96, 256), load_weights_from_unet: bool = True,
                  len_t: int = 1):
        """controlnet.py:1437-1507: copy conv_in / time_embedding / down_blocks / mid_block, zero the zero-convs."""
        c = unet.config
        m = cls(in_channels=c.in_channels, flip_sin_to_cos=c.flip_sin_to_cos, freq_shift=c.freq_shift,
                down_block_types=c.down_block_types, only_cross_attention=c.only_cross_attention,
                block_out_channels=c.block_out_channels, layers_per_block=c.layers_per_block,
                downsample_padding=c.downsample_padding, mid_block_scale_factor=c.mid_block_scale_factor,
                act_fn=c.act_fn, norm_num_groups=c.norm_num_groups, norm_eps=c.norm_eps,
                cross_attention_dim=c.cross_attention_dim, attention_head_dim=c.attention_head_dim,
                use_linear_projection=c.use_linear_projection, upcast_attention=c.upcast_attention,
                resnet_time_scale_shift=c.resnet_time_scale_shift, len_t=1)
        if load_weights_from_unet:
            src = unet.state_dict()
            own = m.state_dict()
            pick = {k: v for k, v in src.items()
                    if k.startswith(("conv_in.", "time_embedding.", "down_blocks.", "mid_block."))}
            own.update(pick)
            m.load_state_dict(own)
        return m.to(unet.device)

    def _program(self, B, H, W, L, scale: float = 1.0):
        key = (B, H, W, L, scale)
        if key in self._progs:
            return self._progs[key]
        net, ws, cfg, dev = self.finalize(), self._ws, self.net_cfg, self._net.device
        P = {"prog": ops.Program()}
        prog = P["prog"]
        P["t"] = torch.zeros(B, device=dev, dtype=torch.float32)
        P["ehs"] = torch.zeros(B * L, cfg.cross_attention_dim, device=dev, dtype=torch.float16)
        cp = pad_channels(cfg.in_channels)
        P["x"] = Act(torch.zeros(B * H * W, cp, device=dev, dtype=torch.float16), B, H, W, cp)
        tproj = net.rec_temb(prog, ws, P["t"], B)
        kv = net.rec_kv(prog, ws, P["ehs"], B, L)
        skips, mid = net.rec_encoder(prog, ws, P["x"], tproj, kv, L)
        P["raw_down"], P["raw_mid"] = skips, mid
        P["down"], P["mid"] = net.rec_exchange(prog, ws, skips, mid, [None] * len(skips), None, scale=scale)
        self._progs[key] = P
        return P

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale: float = 1.0,
                class_labels=None, timestep_cond=None, attention_mask=None, added_cond_kwargs=None,
                cross_attention_kwargs=None, return_dict: bool = True):
        """`sample` is ignored exactly as in the reference (controlnet.py:1716-1720)."""
        self._reject(class_labels=class_labels, timestep_cond=timestep_cond, attention_mask=attention_mask,
                     added_cond_kwargs=added_cond_kwargs)
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs are not supported")
        B, _, H, W = controlnet_cond.shape
        L = encoder_hidden_states.shape[1]
        P = self._program(B, H, W, L, float(conditioning_scale))      # folded into the 13 zero-convs (:1774-1775)
        P["t"].copy_(self._timesteps(timestep, B, P["t"].device))
        P["ehs"].copy_(encoder_hidden_states.reshape(B * L, -1))
        c = controlnet_cond if controlnet_cond.dtype in (torch.float16, torch.float32) else controlnet_cond.float()
        ops.to_nhwc(None, c, P["x"].t, P["x"].C)
        P["prog"].run()
        return ([d.nchw() for d in P["down"]], P["mid"].nchw(), tuple(s.nchw() for s in P["raw_down"]),
                P["raw_mid"].nchw())


class AttributeDecoderModel(_NetModule):
    """Attribute-stream decoder + 13 zero-convs on the RGB stream's raw encoder features
    (reference: models/controlnet.py:1781; forward :2342-2527)."""
    _kind = "attr_dec"

    def __init__(self, out_channels: int = 4, flip_sin_to_cos=True, freq_shift=0,
                 up_block_types=("UpResBlock2D", "CrossAttnUpResBlock2D", "CrossAttnUpResBlock2D",
                                 "CrossAttnUpResBlock2D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, dropout=0.0, act_fn="silu", norm_num_groups=32,
                 norm_eps=1e-5, cross_attention_dim=1280, transformer_layers_per_block=1, attention_head_dim=8,
                 num_attention_heads=None, dual_cross_attention=False, use_linear_projection=False,
                 class_embed_type=None, addition_embed_type=None, num_class_embeds=None, upcast_attention=False,
                 resnet_time_scale_shift="default", len_t: int = 1, _init_weights: bool = True, **unused):
        cfg_kwargs = dict(locals())
        for k in ("self", "__class__", "_init_weights", "unused"):
            cfg_kwargs.pop(k, None)
        super().__init__()
        cfg_kwargs["up_block_types"] = tuple(up_block_types)
        cfg_kwargs["block_out_channels"] = tuple(block_out_channels)
        nc = _net_config(0, out_channels, block_out_channels, layers_per_block, attention_head_dim, cross_attention_dim,
                         norm_num_groups, norm_eps, _SD_DOWN, up_block_types, only_cross_attention=only_cross_attention,
                         dual_cross_attention=dual_cross_attention, use_linear_projection=use_linear_projection,
                         class_embed_type=class_embed_type, addition_embed_type=addition_embed_type,
                         num_class_embeds=num_class_embeds, upcast_attention=upcast_attention,
                         resnet_time_scale_shift=resnet_time_scale_shift, act_fn=act_fn,
                         mid_block_scale_factor=mid_block_scale_factor,
                         transformer_layers_per_block=transformer_layers_per_block)
        self.len_t = len_t
        self.num_upsamplers = len(block_out_channels) - 1
        self._init_common(cfg_kwargs, nc, _init_weights)

    @classmethod
    def from_unet(cls, unet: UNet2DConditionModel, load_weights_from_unet: bool = True, len_t: int = 1):
        """controlnet.py:2115-2192: copy conv_out / time_embedding / up_blocks; up_block_types come from the UNet."""
        c = unet.config
        m = cls(out_channels=c.out_channels, flip_sin_to_cos=c.flip_sin_to_cos, freq_shift=c.freq_shift,
                up_block_types=c.up_block_types, only_cross_attention=c.only_cross_attention,
                block_out_channels=c.block_out_channels, layers_per_block=c.layers_per_block,
                downsample_padding=c.downsample_padding, mid_block_scale_factor=c.mid_block_scale_factor,
                act_fn=c.act_fn, norm_num_groups=c.norm_num_groups, norm_eps=c.norm_eps,
                cross_attention_dim=c.cross_attention_dim, attention_head_dim=c.attention_head_dim,
                use_linear_projection=c.use_linear_projection, upcast_attention=c.upcast_attention,
                resnet_time_scale_shift=c.resnet_time_scale_shift, len_t=1)
        if load_weights_from_unet:
            own = m.state_dict()
            own.update({k: v for k, v in unet.state_dict().items()
                        if k.startswith(("conv_out.", "time_embedding.", "up_blocks."))})
            m.load_state_dict(own)
        return m.to(unet.device)

    def _program(self, B, H, W, L, srcs):
        """srcs: dict of Act inputs.  Programs are keyed by the input pointers so the steady-state loop (inputs are
        the sibling modules' static buffers) replays with zero copies."""
        key = (B, H, W, L, tuple(a.t.data_ptr() for a in srcs["skipA"] + srcs["skipU"] + [srcs["midA"], srcs["midU"]]
                                 + list(srcs.get("up") or [])))
        if key in self._progs:
            return self._progs[key]
        net, ws, cfg, dev = self.finalize(), self._ws, self.net_cfg, self._net.device
        P = {"prog": ops.Program(), "srcs": srcs}
        prog = P["prog"]
        P["t"] = torch.zeros(B, device=dev, dtype=torch.float32)
        P["ehs"] = torch.zeros(B * L, cfg.cross_attention_dim, device=dev, dtype=torch.float16)
        P["out"] = torch.zeros(B, cfg.out_channels, H, W, device=dev, dtype=torch.float32)
        tproj = net.rec_temb(prog, ws, P["t"], B)
        kv = net.rec_kv(prog, ws, P["ehs"], B, L)
        skips, mid = net.rec_exchange(prog, ws, srcs["skipU"], srcs["midU"], srcs["skipA"], srcs["midA"])
        net.rec_decoder(prog, ws, mid, skips, tproj, kv, L, out_nchw=P["out"], up_additional=srcs.get("up"))
        self._progs[key] = P
        return P

    def _slot(self, name, like: torch.Tensor) -> Act:
        slots = self.__dict__.setdefault("_slots", {})
        B, Cn, H, W = like.shape
        k = (name, B, Cn, H, W)
        if k not in slots:
            slots[k] = Act(torch.zeros(B * H * W, Cn, device=self._net.device, dtype=torch.float16), B, H, W, Cn)
        return slots[k]

    def _ingest(self, name, t: torch.Tensor) -> Act:
        a = _as_act(t)
        if a is not None:
            return a
        s = self._slot(name, t)
        s.nchw().copy_(t)
        return s

    @torch.no_grad()
    def forward(self, sample, down_block_res_samples, timestep, encoder_hidden_states,
                down_block_additional_residuals=None, up_block_additional_residuals=None,
                mid_block_additional_residual=None, class_labels=None, timestep_cond=None, attention_mask=None,
                added_cond_kwargs=None, cross_attention_kwargs=None, return_dict: bool = True):
        self._reject(class_labels=class_labels, timestep_cond=timestep_cond, attention_mask=attention_mask,
                     added_cond_kwargs=added_cond_kwargs)
        if cross_attention_kwargs:
            raise NotImplementedError("cross_attention_kwargs are not supported")
        if down_block_additional_residuals is None or mid_block_additional_residual is None:
            raise ValueError("AttributeDecoderModel needs the RGB stream's raw skips and mid (controlnet.py:2476 "
                             "dereferences mid_block_additional_residual unconditionally)")
        n_layers = len(self.net_cfg.block_out_channels) * (self.net_cfg.layers_per_block + 1)
        if self.net_cfg.up_res:
            # UpResBlock2D / CrossAttnUpResBlock2D (the class-default up_block_types): each decoder layer adds
            # up_block_additional_residuals[k] to its output (models/unet_2d_blocks.py:2408,2814) -- SURVEY row a8
            if up_block_additional_residuals is None or len(up_block_additional_residuals) != n_layers:
                raise ValueError(f"up_block_types {_SD_UPRES} need {n_layers} up_block_additional_residuals "
                                 "(one per decoder layer, in layer order)")
        elif up_block_additional_residuals is not None:
            import warnings
            warnings.warn("up_block_additional_residuals is ignored with UpBlock2D / CrossAttnUpBlock2D up blocks -- exactly "
                          "like the reference, whose live forward has the argument commented out "
                          "(models/controlnet.py:2464-2510)", stacklevel=2)
            up_block_additional_residuals = None
        self.finalize(sample.device)
        B, _, h8, w8 = sample.shape
        H, W = down_block_res_samples[0].shape[-2:]
        L = encoder_hidden_states.shape[1]
        srcs = {"skipA": [self._ingest(f"a{i}", t) for i, t in enumerate(down_block_res_samples)],
                "skipU": [self._ingest(f"u{i}", t) for i, t in enumerate(down_block_additional_residuals)],
                "midA": self._ingest("am", sample), "midU": self._ingest("um", mid_block_additional_residual)}
        if up_block_additional_residuals is not None:
            srcs["up"] = [self._ingest(f"x{i}", t) for i, t in enumerate(up_block_additional_residuals)]
        P = self._program(B, H, W, L, srcs)
        P["t"].copy_(self._timesteps(timestep, B, P["t"].device))
        P["ehs"].copy_(encoder_hidden_states.reshape(B * L, -1))
        P["prog"].run()
        out = P["out"]          # fp32 NCHW (the prediction feeds the scheduler update)
        if not return_dict:
            return out
        return UNet2DConditionOutput(sample=out)


def random_init_state_dict(kind: str, cfg: NetConfig, seed: int, device="cuda", zero_conv_std: float = 0.02,
                           dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Random-init weights in the reference's state-dict layout, generated directly on `device` (synthetic
    benchmarks: there is no network for checkpoints).  Linear/conv weights U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like
    torch's default init, norm scales 1 +- 0.1, and the exchange zero-convs N(0, zero_conv_std) instead of the
    reference's zeros (controlnet.py:1360-1415) so that the dual-stream exchange is numerically live."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shp in _param_shapes(kind, cfg).items():
        if name.startswith(("controlnet_", "control_")):
            t = torch.randn(shp, generator=g, device=device, dtype=torch.float32) * zero_conv_std
        elif "norm" in name and len(shp) == 1:
            t = 0.1 * torch.randn(shp, generator=g, device=device, dtype=torch.float32)
            if name.endswith("weight"):
                t += 1.0
        else:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            if len(shp) == 1:                       # bias: fan_in of the matching weight
                wshape = _param_shapes(kind, cfg)[name[:-len("bias")] + "weight"]
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
            k = fan_in ** -0.5
            t = (torch.rand(shp, generator=g, device=device, dtype=torch.float32) * 2 - 1) * k
        sd[name] = t.to(dtype)
    return sd
