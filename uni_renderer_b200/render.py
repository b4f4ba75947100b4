"""Image-to-image rendering calls around the fused loops: VAE encode -> sampling loop -> VAE decode on the B200 path.

Tensor-level mirror of the two calls the shipped eval makes on the reference pipeline (paths relative to
/root/reference; PIL handling, CLIP and the safety checker stay in the reference's Python -- pass `prompt_embeds`):

  forward_rendering  = UniRendererPipeline.mask2image_3mod_albedo   models/pipeline.py:1368-1697
      six `vae.encode(x).latent_dist.sample() * scaling_factor` (:1531-1556), material latents from two numbers
      (:1534-1541), attr28 = cat(masks, material, normal, albedo, spec_light, diff_light, env) (:1583), the image
      stream denoised with t_attr = 0 (:1586-1653), one decode (:1664)
  inverse_rendering  = UniRendererPipeline.image2mask_3mod_albedo   models/pipeline.py:1990-2390
      encode(image), encode(masks) (:2112-2117), six prepare_latents draws (:2119-2188), the 24 attribute channels
      denoised with t_img = 0 (:2207-2312), five decodes (:2335-2349); returns (material latents, normal, albedo,
      spec_light, diff_light, env images) like :2389

B200-first differences (results identical for the same noise): the encodes of one call run as ONE batched program
(6B / 2B images), the five decodes as one 5B-latent program, and the loop is the fused CUDA-graph sampler.
RNG use follows the reference: `generator` feeds ONLY the `prepare_latents` draws (models/pipeline.py:705-719,
2119-2188); the VAE posterior noise of `latent_dist.sample()` is drawn WITHOUT a generator there (:1531-1556,
2113-2116), i.e. from the device's global RNG -- here too, one draw per encoded image in the reference's order
(`posterior_generator` exists for reproducible tests only).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from .pipeline import MASK_CHANNELS, DualStreamSampler
from .vae import AutoencoderKL

ATTR_GROUPS = ("material", "normal", "albedo", "spec_light", "diff_light", "env")      # after the 4 mask channels


class RenderPipeline:
    def __init__(self, sampler: DualStreamSampler, vae: AutoencoderKL, prompt_cache=None):
        """prompt_cache: an optional uni_renderer_b200.text.PromptEmbedCache; with it `prompt_embeds=None` encodes
        `prompt` (default ' ', what every shipped caller passes) once per process."""
        self.sampler, self.vae, self.prompt_cache = sampler, vae, prompt_cache
        self.device = sampler.device
        vae.finalize(self.device)
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1)            # pipeline.py:178

    # -- pieces ------------------------------------------------------------------------------------------------
    def _randn(self, shape, generator):
        gdev = generator.device if generator is not None else self.device
        return torch.randn(shape, generator=generator, device=gdev, dtype=torch.float32).to(self.device)

    @torch.no_grad()
    def encode_images(self, images: Sequence[torch.Tensor], generator: Optional[torch.Generator] = None,
                      sample: bool = True) -> Tuple[torch.Tensor, ...]:
        """`vae.encode(x).latent_dist.sample() * scaling_factor` for several [B, 3, H, W] images in [-1, 1] as one
        batched encoder program; the posterior noise of each image is drawn separately, in order (generator None =
        the device's global RNG, which is what the reference's `latent_dist.sample()` uses)."""
        from . import ops
        n, B = len(images), images[0].shape[0]
        x = torch.cat([i.to(self.device) for i in images], 0)
        moments = self.vae.encode(x).latent_dist.parameters                          # [n*B, 8, h, w] fp32
        noise = None
        if sample:
            shape = (B, moments.shape[1] // 2) + tuple(moments.shape[2:])
            noise = torch.cat([self._randn(shape, generator) for _ in range(n)], 0)
        out = torch.empty(n * B, moments.shape[1] // 2, *moments.shape[2:], device=self.device, dtype=torch.float32)
        ops.gaussian_sample(None, moments, noise, out, scale=float(self.vae.config.scaling_factor))
        return tuple(out[i * B:(i + 1) * B] for i in range(n))

    @torch.no_grad()
    def decode_latents(self, latents: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        """`vae.decode(z / scaling_factor, return_dict=False)[0]` for several [B, 4, h, w] latents as one program."""
        n, B = len(latents), latents[0].shape[0]
        z = torch.cat([l.to(self.device, torch.float32) for l in latents], 0) / float(self.vae.config.scaling_factor)
        img = self.vae.decode(z, return_dict=False)[0]
        return tuple(img[i * B:(i + 1) * B] for i in range(n))

    # -- the two shipped calls -----------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_rendering(self, material_num, normal_image, albedo_image, spec_light_image, diff_light_image, env_image,
                          masks_image, prompt_embeds=None, num_inference_steps: int = 50, guidance_scale: float = 0.0,
                          generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None,
                          output_type: str = "pt", scheduler: Optional[str] = None, prompt: str = " ",
                          negative_prompt_embeds: Optional[torch.Tensor] = None,
                          posterior_generator: Optional[torch.Generator] = None):
        """attributes -> RGB (mask2image_3mod_albedo).  Images are [B, 3, H, W] tensors in [-1, 1]; material_num is
        (metallic, roughness); returns the decoded image [B, 3, H, W] (or the latents for output_type="latent")."""
        B = normal_image.shape[0]
        # encode order of the reference: normal, albedo, spec_light, diff_light, env, masks (:1531-1556)
        l_normal, l_albedo, l_spec, l_diff, l_env, l_masks = self.encode_images(
            [normal_image, albedo_image, spec_light_image, diff_light_image, env_image, masks_image],
            posterior_generator)
        metallic, roughness = float(material_num[0]), float(material_num[1])
        l_material = torch.empty_like(l_normal)
        l_material[:, :2] = metallic * 2 - 1.0                                       # :1534-1541
        l_material[:, 2:] = roughness * 2 - 1.0
        attr28 = torch.cat((l_masks, l_material, l_normal, l_albedo, l_spec, l_diff, l_env), 1)      # :1583
        if latents is None:
            latents = self._randn((B, 4) + tuple(l_normal.shape[2:]), generator)     # prepare_latents, :705-719
        ehs = self._embeds(prompt_embeds, B, prompt)
        lat = self.sampler.forward_render(latents.to(self.device, torch.float32), attr28, ehs, num_inference_steps,
                                          guidance_scale, scheduler, negative_prompt_embeds)
        if output_type == "latent":
            return lat
        return self.decode_latents([lat])[0]

    @torch.no_grad()
    def inverse_rendering(self, image, masks, prompt_embeds=None, num_inference_steps: int = 50,
                          guidance_scale: float = 0.0, generator: Optional[torch.Generator] = None,
                          latents: Optional[Sequence[torch.Tensor]] = None, scheduler: Optional[str] = None,
                          prompt: str = " ", negative_prompt_embeds: Optional[torch.Tensor] = None,
                          posterior_generator: Optional[torch.Generator] = None):
        """RGB -> attributes (image2mask_3mod_albedo).  Returns (material_latents, normal, albedo, spec_light,
        diff_light, env) with the five images decoded to [B, 3, H, W] in [-1, 1] (:2389)."""
        B = image.shape[0]
        l_img, l_masks = self.encode_images([image, masks], posterior_generator)     # :2112-2117
        if latents is None:                                                          # six draws, :2119-2188
            latents = [self._randn(tuple(l_img.shape), generator) for _ in ATTR_GROUPS]
        if len(latents) != len(ATTR_GROUPS):
            raise ValueError(f"need {len(ATTR_GROUPS)} attribute latents ({ATTR_GROUPS})")
        attr28 = torch.cat([l_masks] + [l.to(self.device, torch.float32) for l in latents], 1)
        ehs = self._embeds(prompt_embeds, B, prompt)
        attr24 = self.sampler.inverse_render(l_img, attr28, ehs, num_inference_steps, guidance_scale, scheduler,
                                             negative_prompt_embeds)
        groups = [attr24[:, 4 * i:4 * i + 4] for i in range(len(ATTR_GROUPS))]
        decoded = self.decode_latents(groups[1:])                                    # material stays latent (:2331)
        return (groups[0],) + decoded

    def _embeds(self, prompt_embeds: Optional[torch.Tensor], B: int, prompt: str = " ") -> torch.Tensor:
        if prompt_embeds is None:
            if self.prompt_cache is None:
                raise ValueError("pass prompt_embeds, or construct RenderPipeline with a PromptEmbedCache")
            prompt_embeds = self.prompt_cache.encode(prompt)
        if prompt_embeds.shape[0] == 1 and B > 1:
            prompt_embeds = prompt_embeds.repeat(B, 1, 1)                            # :2109
        if prompt_embeds.shape[0] != B:
            raise ValueError(f"prompt_embeds batch {prompt_embeds.shape[0]} != image batch {B}")
        return prompt_embeds


__all__ = ["RenderPipeline", "ATTR_GROUPS", "MASK_CHANNELS"]
