"""`UniRendererPipeline`: the call surface of the reference's models/pipeline.py that the shipped entry points use, on the
B200 path.

eval/test_real.py:470-553 and eval/test_app.py:117-225 build the pipeline from components and call

    pipeline = UniRendererPipeline.from_pretrained(path, vae=vae, text_encoder=..., tokenizer=..., unet=unet,
                                                   controlnet=controlnet, controldec=controldec, safety_checker=None)
    pipeline.scheduler_img = UniPCMultistepScheduler.from_config(pipeline.scheduler.config)    # ... one per stream
    material, normal, albedo, spec, diff, env = pipeline.real_image2mask_3mod_albedo(
        ' ', image, mask, guidance_scale=0.0, height=512, width=512, num_inference_steps=20, generator=g)

This class keeps those names, keyword arguments and return structures (models/pipeline.py:124-215 constructor and
toggles, :1368 mask2image_3mod_albedo, :1990 image2mask_3mod_albedo, :2391 real_image2mask_3mod_albedo) and runs the
body on `RenderPipeline` (batched VAE programs + the fused CUDA-graph loops).  What stays host Python exactly as in the
reference: PIL / numpy pre- and post-processing (diffusers' VaeImageProcessor semantics: RGB, Lanczos resize, [0, 1]
-> [-1, 1]; `(x / 2 + 0.5).clamp(0, 1)` back) and the text encoder, which runs once per distinct prompt
(`PromptEmbedCache`).  The per-stream scheduler attributes the callers assign are honoured through their `.config`: the
timestep / coefficient tables of the fused loops are rebuilt from `beta_*`, `beta_schedule`, `timestep_spacing`,
`steps_offset`, `prediction_type` (and `solver_type` for UniPC); a scheduler class other than DDIM / UniPC, or an option
the tables do not reproduce (`clip_sample`, `thresholding`, `solver_order != 2`, Karras sigmas ...), raises.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence, Union

import torch

from .pipeline import DualStreamSampler
from .render import RenderPipeline
from .scheduler import DDIMSchedule, UniPCSchedule, load_scheduler_config
from .text import PromptEmbedCache

_STREAM_SCHEDULERS = ("scheduler_img", "scheduler_attr", "scheduler_material", "scheduler_albedo", "scheduler_normal",
                      "scheduler_spec_light", "scheduler_diff_light", "scheduler_env")


def _scheduler_kind(obj) -> str:
    name = type(obj).__name__
    if "UniPC" in name:
        return "unipc"
    if "DDIM" in name:
        return "ddim"
    raise NotImplementedError(f"scheduler {name} is not wired on the B200 path (DDIMScheduler and UniPCMultistepScheduler are)")


def _prediction_type(obj) -> str:
    cfg = getattr(obj, "config", None)
    pt = getattr(cfg, "prediction_type", None) if cfg is not None else None
    if pt is None and isinstance(cfg, dict):
        pt = cfg.get("prediction_type")
    return pt or getattr(obj, "prediction_type", None) or "epsilon"


def _schedule_from(obj, kind: str):
    """The coefficient-table object for an assigned scheduler: built from its `.config` when it has one (diffusers
    objects, the holders of scheduler.py), else the SD-1.x defaults with the object's prediction type."""
    cls = UniPCSchedule if kind == "unipc" else DDIMSchedule
    cfg = getattr(obj, "config", None)
    has_keys = cfg is not None and any((k in cfg) if isinstance(cfg, dict) else hasattr(cfg, k)
                                       for k in ("beta_start", "timestep_spacing", "num_train_timesteps"))
    if has_keys:
        return cls.from_config(cfg, prediction_type=_prediction_type(obj))
    return cls(prediction_type=_prediction_type(obj))


def preprocess_image(image, height: int, width: int) -> torch.Tensor:
    """VaeImageProcessor(do_convert_rgb=True).preprocess as models/pipeline.py:674-686 uses it: PIL image(s) -> RGB ->
    Lanczos resize to (width, height) -> float in [0, 1] -> [-1, 1], NCHW.  Tensors / arrays in [0, 1] are normalised
    the same way; tensors that already contain negative values are taken as [-1, 1] (diffusers warns and does that)."""
    import numpy as np
    if isinstance(image, torch.Tensor):
        x = image.float()
        if x.dim() == 3:
            x = x[None]
        return x if float(x.min()) < 0 else x * 2.0 - 1.0
    if isinstance(image, np.ndarray):
        x = torch.from_numpy(image).float()
        if x.dim() == 3:
            x = x[None]
        x = x.permute(0, 3, 1, 2)
        return x if float(x.min()) < 0 else x * 2.0 - 1.0
    from PIL import Image
    imgs = list(image) if isinstance(image, (list, tuple)) else [image]
    out = []
    for im in imgs:
        im = im.convert("RGB").resize((width, height), resample=Image.LANCZOS)
        out.append(torch.from_numpy(np.asarray(im, dtype=np.float32) / 255.0).permute(2, 0, 1))
    return torch.stack(out, 0) * 2.0 - 1.0


def postprocess_image(x: torch.Tensor, output_type: str = "pil"):
    """VaeImageProcessor.postprocess with do_denormalize: [-1, 1] -> [0, 1] -> "pt" tensor / "np" NHWC array / "pil"."""
    x = (x.float() / 2 + 0.5).clamp(0, 1)
    if output_type == "pt":
        return x
    arr = x.cpu().permute(0, 2, 3, 1).numpy()
    if output_type == "np":
        return arr
    if output_type != "pil":
        raise ValueError(f"output_type {output_type!r} is not supported (pil / np / pt / latent)")
    from PIL import Image
    return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]


class UniRendererPipeline:
    def __init__(self, vae, text_encoder=None, tokenizer=None, unet=None, controlnet=None, controldec=None,
                 scheduler=None, safety_checker=None, feature_extractor=None, image_encoder=None,
                 requires_safety_checker: bool = False):
        if unet is None or controlnet is None or controldec is None or vae is None:
            raise ValueError("vae, unet, controlnet and controldec are required")
        if safety_checker is not None:
            raise NotImplementedError("pass safety_checker=None (every shipped caller does, eval/test_real.py:477)")
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.controlnet, self.controldec = unet, controlnet, controldec
        self.scheduler = scheduler
        for name in _STREAM_SCHEDULERS:                     # the callers overwrite these (eval/test_real.py:485-493)
            setattr(self, name, scheduler)
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1)            # models/pipeline.py:178
        self._render: Optional[RenderPipeline] = None
        self._sampler: Optional[DualStreamSampler] = None
        self._render_key = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, **kwargs):
        """The shipped callers pass every model component as a keyword (eval/test_real.py:470-482); the directory is only
        consulted for components that were not passed, and only for the classes of this package."""
        import os
        from .models import AttributeDecoderModel, AttributeEncoderModel, UNet2DConditionModel
        from .vae import AutoencoderKL
        comps = {k: kwargs.pop(k, None) for k in ("vae", "text_encoder", "tokenizer", "unet", "controlnet", "controldec",
                                                   "scheduler", "safety_checker", "feature_extractor", "image_encoder")}
        loaders = {"vae": AutoencoderKL, "unet": UNet2DConditionModel, "controlnet": AttributeEncoderModel,
                   "controldec": AttributeDecoderModel}
        for name, klass in loaders.items():
            if comps[name] is None and pretrained_model_name_or_path is not None \
                    and os.path.isdir(os.path.join(pretrained_model_name_or_path, name)):
                comps[name] = klass.from_pretrained(pretrained_model_name_or_path, subfolder=name)
        # `pipeline.scheduler.config` is what the eval derives every per-stream scheduler from (test_real.py:485-493)
        if comps["scheduler"] is None and pretrained_model_name_or_path is not None:
            sdir = os.path.join(pretrained_model_name_or_path, "scheduler")
            if os.path.isfile(os.path.join(sdir, "scheduler_config.json")):
                comps["scheduler"] = load_scheduler_config(sdir)
        return cls(**comps)

    # -- toggles the reference's callers invoke ------------------------------------------------------------------
    def to(self, device=None, dtype=None):
        for m in (self.vae, self.unet, self.controlnet, self.controldec):
            m.to(device)
        if self.text_encoder is not None:
            self.text_encoder.to(device)
        self._render = self._sampler = None          # packed weights / recorded programs follow the modules
        return self

    def set_progress_bar_config(self, **kwargs):
        return None

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None

    def enable_vae_slicing(self):
        self.vae.enable_slicing()

    def disable_vae_slicing(self):
        self.vae.disable_slicing()

    def enable_vae_tiling(self):
        self.vae.enable_tiling()

    def disable_vae_tiling(self):
        self.vae.disable_tiling()

    # -- plumbing ------------------------------------------------------------------------------------------------
    def _runner(self, sched_names: Sequence[str]):
        """(RenderPipeline, scheduler kind) for the streams a call updates; all of them must be of one kind."""
        scheds = [getattr(self, n) for n in sched_names]
        if any(s is None for s in scheds):
            raise ValueError(f"assign {', '.join(sched_names)} before sampling (eval/test_real.py:485-493)")
        kinds = {_scheduler_kind(s) for s in scheds}
        if len(kinds) != 1:
            raise NotImplementedError("the per-stream schedulers of one call must share kind and prediction_type")
        kind = kinds.pop()
        tables = [_schedule_from(s, kind) for s in scheds]
        if len({t.signature() for t in tables}) != 1:
            raise NotImplementedError("the per-stream schedulers of one call must share kind, prediction_type and "
                                      "configuration (the fused loop walks ONE timestep table)")
        ptype = tables[0].prediction_type
        if self._render is None or self._render_key != ptype:
            if self._sampler is None or self._render_key != ptype:
                self._sampler = DualStreamSampler(self.unet, self.controlnet, self.controldec, prediction_type=ptype)
            cache = None
            if self.text_encoder is not None and self.tokenizer is not None:
                cache = PromptEmbedCache(self.tokenizer, self.text_encoder, device=self._sampler.device)
            self._render = RenderPipeline(self._sampler, self.vae, prompt_cache=cache)
            self._render_key = ptype
        self._sampler.set_schedules(**{"unipc" if kind == "unipc" else "ddim": tables[0]})
        return self._render, kind

    def _embeds(self, rp: RenderPipeline, prompt, prompt_embeds, negative_prompt, negative_prompt_embeds, guidance_scale):
        if prompt_embeds is None:
            if rp.prompt_cache is None:
                raise ValueError("pass prompt_embeds, or construct the pipeline with text_encoder and tokenizer")
            prompt_embeds = rp.prompt_cache.encode(prompt if prompt is not None else " ")
        if guidance_scale not in (0, 0.0, None) and negative_prompt_embeds is None:
            if rp.prompt_cache is None:
                raise ValueError("guidance_scale != 0 needs negative_prompt_embeds (or a text encoder)")
            negative_prompt_embeds = rp.prompt_cache.encode(negative_prompt if negative_prompt is not None else "")
        return prompt_embeds, negative_prompt_embeds

    def _hw(self, height, width):
        size = getattr(self.unet.config, "sample_size", None)
        height = height or (size * self.vae_scale_factor if size else 512)
        width = width or (size * self.vae_scale_factor if size else 512)
        return int(height), int(width)

    # -- the calls -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def real_image2mask_3mod_albedo(self, prompt: Union[str, List[str], None] = None, image=None, masks=None,
                                    height: Optional[int] = None, width: Optional[int] = None,
                                    num_inference_steps: int = 50, timesteps=None, guidance_scale: float = 7.5,
                                    negative_prompt=None, num_images_per_prompt: int = 1, eta: float = 0.0,
                                    generator: Optional[torch.Generator] = None, latents=None, prompt_embeds=None,
                                    negative_prompt_embeds=None, ip_adapter_image=None, output_type: str = "pil",
                                    return_dict: bool = True, cross_attention_kwargs=None,
                                    controlnet_conditioning_scale: float = 1.0, **kwargs: Any):
        """RGB image + object mask -> (material latents [B, 4, h, w], normal, albedo, spec_light, diff_light, env) with
        the five maps post-processed to `output_type` (models/pipeline.py:2391-2808; note the reference's default
        guidance_scale of 7.5 is kept in the signature -- the shipped callers pass 0.0)."""
        if timesteps is not None or ip_adapter_image is not None or cross_attention_kwargs or num_images_per_prompt != 1 \
                or float(controlnet_conditioning_scale) != 1.0:
            raise NotImplementedError("custom timesteps / ip_adapter_image / cross_attention_kwargs / "
                                      "num_images_per_prompt / controlnet_conditioning_scale are not supported here")
        if output_type == "latent":
            raise ValueError("output_type='latent' is rejected by the reference too (models/pipeline.py:2779)")
        rp, kind = self._runner(("scheduler_material", "scheduler_normal", "scheduler_albedo", "scheduler_spec_light",
                                 "scheduler_diff_light", "scheduler_env"))
        height, width = self._hw(height, width)
        x, m = preprocess_image(image, height, width), preprocess_image(masks, height, width)
        pe, ne = self._embeds(rp, prompt, prompt_embeds, negative_prompt, negative_prompt_embeds, guidance_scale)
        out = rp.inverse_rendering(x, m, pe, num_inference_steps, guidance_scale, generator,
                                   latents=latents, scheduler=kind, negative_prompt_embeds=ne)
        return (out[0],) + tuple(postprocess_image(o, output_type) for o in out[1:])

    image2mask_3mod_albedo = real_image2mask_3mod_albedo       # same body; the reference variant takes tensors (:1990)

    @torch.no_grad()
    def mask2image_3mod_albedo(self, prompt=None, material_num=None, normal_image=None, albedo_image=None,
                               spec_light_image=None, diff_light_image=None, env_image=None, masks_image=None,
                               re_rendering: bool = False, height: Optional[int] = None, width: Optional[int] = None,
                               num_inference_steps: int = 50, timesteps=None, guidance_scale: float = 7.5,
                               negative_prompt=None, num_images_per_prompt: int = 1, eta: float = 0.0,
                               generator: Optional[torch.Generator] = None, latents=None, prompt_embeds=None,
                               negative_prompt_embeds=None, ip_adapter_image=None, output_type: str = "pil",
                               return_dict: bool = True, cross_attention_kwargs=None,
                               controlnet_conditioning_scale: float = 1.0, **kwargs: Any):
        """Attribute maps + (metallic, roughness) -> rendered RGB image(s) (models/pipeline.py:1368-1697)."""
        if timesteps is not None or ip_adapter_image is not None or cross_attention_kwargs or num_images_per_prompt != 1 \
                or float(controlnet_conditioning_scale) != 1.0:
            raise NotImplementedError("custom timesteps / ip_adapter_image / cross_attention_kwargs / "
                                      "num_images_per_prompt / controlnet_conditioning_scale are not supported here")
        rp, kind = self._runner(("scheduler_img",))
        height, width = self._hw(height, width)
        imgs = [preprocess_image(i, height, width) for i in (normal_image, albedo_image, spec_light_image,
                                                             diff_light_image, env_image, masks_image)]
        pe, ne = self._embeds(rp, prompt, prompt_embeds, negative_prompt, negative_prompt_embeds, guidance_scale)
        out = rp.forward_rendering(material_num, *imgs, pe, num_inference_steps, guidance_scale, generator, latents,
                                   "latent" if output_type == "latent" else "pt", kind, negative_prompt_embeds=ne)
        return out if output_type == "latent" else postprocess_image(out, output_type)
