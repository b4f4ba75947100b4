"""`UniRendererPipeline` (the models/pipeline.py call surface the shipped entry points use) on the CPU emulator: PIL in,
PIL out, same names / keyword arguments / return structure as eval/test_real.py:531-537 and eval/test_app.py:211-217
call it, and the same numbers as the tensor-level RenderPipeline underneath."""
from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle import uni_oracle as uo
from oracle import vae_oracle as vo
from tests import cpu_ops_emulator as emu
from uni_renderer_b200 import models as M
from uni_renderer_b200 import unipipeline as UP
from uni_renderer_b200 import vae as V
from uni_renderer_b200.engine import NetConfig, StreamNet, Workspace

TINY = dict(block_out_channels=(32, 64, 128, 128), attention_head_dim=4, cross_attention_dim=48, norm_num_groups=8)


class UniPCMultistepScheduler:                # stand-in with the diffusers class name: the pipeline keys on the kind
    class config:
        prediction_type = "epsilon"


class EulerDiscreteScheduler:
    class config:
        prediction_type = "epsilon"


def _build(monkeypatch):
    emu.install(monkeypatch)

    def net_finalize(self, device=None):
        if self._net is None:
            self._net = StreamNet(self._kind, self.net_cfg, dict(self.state_dict()), "cpu")
            self._ws = Workspace("cpu")
        return self._net

    def vae_finalize(self, device=None):
        if self._net is None:
            net = object.__new__(V.VaeNet)
            net.cfg, net.device, net.w = self.vae_cfg, torch.device("cpu"), {}
            net._pack(V.convert_deprecated_attention_keys(dict(self.state_dict())))
            self._net, self._ws = net, Workspace("cpu")
        return self._net
    monkeypatch.setattr(M._NetModule, "finalize", net_finalize)
    monkeypatch.setattr(V.AutoencoderKL, "finalize", vae_finalize)
    monkeypatch.setattr(V.AutoencoderKL, "use_graph", False)
    base = uo.TINY
    cfgs_o = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [uo.random_state_dict(k, c, s) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs_o, (11, 12, 13))]
    unet = M.UNet2DConditionModel(in_channels=4, out_channels=4, _init_weights=False, **TINY)
    enc = M.AttributeEncoderModel(in_channels=28, _init_weights=False, **TINY)
    dec = M.AttributeDecoderModel(out_channels=28, up_block_types=M._SD_UP, _init_weights=False, **TINY)
    for m, sd in zip((unet, enc, dec), sds):
        m.load_state_dict(sd)
    vae = V.AutoencoderKL(block_out_channels=(32, 64, 64), down_block_types=(V._DOWN,) * 3, up_block_types=(V._UP,) * 3,
                          layers_per_block=2, norm_num_groups=8)
    vae.load_state_dict(vo.random_state_dict(vo.TINY_VAE, 5))
    pipe = UP.UniRendererPipeline.from_pretrained(None, vae=vae, text_encoder=None, tokenizer=None, unet=unet,
                                                  controlnet=enc, controldec=dec, safety_checker=None)
    nb = NetConfig(block_out_channels=base.block_out_channels, num_heads=base.num_heads,
                   cross_attention_dim=base.cross_attention_dim, norm_num_groups=base.norm_num_groups)
    cfgs = (replace(nb), replace(nb, in_channels=28), replace(nb, out_channels=28))
    pipe = pipe.to("cpu")
    pipe._sampler, pipe._render_key = emu.cpu_sampler(sds, cfgs), "epsilon"       # test only: no CUDA gate
    return pipe


def _pil(seed, size=40):
    from PIL import Image
    rng = np.random.default_rng(seed)
    return Image.fromarray(rng.integers(0, 256, (size, size, 3), dtype=np.uint8))


def test_real_image2mask_call_as_the_shipped_eval_makes_it(monkeypatch):
    from PIL import Image
    pipe = _build(monkeypatch)
    with pytest.raises(ValueError):                          # the callers assign the per-stream schedulers first
        pipe.real_image2mask_3mod_albedo(" ", _pil(1), _pil(2), guidance_scale=0.0, height=32, width=32)
    for n in UP._STREAM_SCHEDULERS:
        setattr(pipe, n, UniPCMultistepScheduler())
    pipe.scheduler = None
    pipe.set_progress_bar_config(disable=True)
    pipe.enable_xformers_memory_efficient_attention()
    ehs = torch.randn(1, 7, 48, generator=torch.Generator().manual_seed(3)).half()
    g = torch.Generator().manual_seed(42)
    torch.manual_seed(1)        # the VAE posterior noise comes from the global RNG, as in the reference
    material, normal, albedo, spec, diff, env = pipe.real_image2mask_3mod_albedo(
        " ", _pil(1), _pil(2), guidance_scale=0.0, height=32, width=32, num_inference_steps=3, generator=g,
        prompt_embeds=ehs)
    assert material.shape == (1, 4, 8, 8) and torch.isfinite(material).all()
    for out in (normal, albedo, spec, diff, env):
        assert isinstance(out, list) and len(out) == 1 and isinstance(out[0], Image.Image) and out[0].size == (32, 32)
    # the same numbers as the tensor-level call underneath, with the documented pre / post-processing
    x, m = UP.preprocess_image(_pil(1), 32, 32), UP.preprocess_image(_pil(2), 32, 32)
    assert x.shape == (1, 3, 32, 32) and -1.0 <= float(x.min()) and float(x.max()) <= 1.0
    g.manual_seed(42)
    torch.manual_seed(1)
    ref = pipe._render.inverse_rendering(x, m, ehs, 3, 0.0, g, scheduler="unipc")
    assert torch.equal(ref[0], material)
    assert np.array_equal(np.asarray(UP.postprocess_image(ref[2], "pil")[0]), np.asarray(albedo[0]))
    # output_type variants and the tensor-input alias
    g.manual_seed(42)
    torch.manual_seed(1)
    out_pt = pipe.image2mask_3mod_albedo(" ", (x + 1) / 2, (m + 1) / 2, guidance_scale=0.0, num_inference_steps=3,
                                         generator=g, prompt_embeds=ehs, output_type="pt", height=32, width=32)
    assert torch.equal(out_pt[0], material) and out_pt[1].shape == (1, 3, 32, 32) and float(out_pt[1].min()) >= 0.0
    with pytest.raises(ValueError):
        pipe.real_image2mask_3mod_albedo(" ", _pil(1), _pil(2), guidance_scale=0.0, height=32, width=32,
                                         prompt_embeds=ehs, output_type="latent")
    with pytest.raises(ValueError):                          # no text encoder and no embeddings
        pipe.real_image2mask_3mod_albedo(" ", _pil(1), _pil(2), guidance_scale=0.0, height=32, width=32)
    pipe.scheduler_env = EulerDiscreteScheduler()
    with pytest.raises(NotImplementedError):
        pipe.real_image2mask_3mod_albedo(" ", _pil(1), _pil(2), guidance_scale=0.0, height=32, width=32, prompt_embeds=ehs)


def test_mask2image_call(monkeypatch):
    from PIL import Image
    pipe = _build(monkeypatch)
    pipe.scheduler_img = UniPCMultistepScheduler()
    ehs = torch.randn(1, 7, 48, generator=torch.Generator().manual_seed(3)).half()
    g = torch.Generator().manual_seed(7)
    imgs = [_pil(10 + i) for i in range(6)]
    out = pipe.mask2image_3mod_albedo(" ", np.array([0.3, 0.8]), *imgs, height=32, width=32, num_inference_steps=2,
                                      guidance_scale=0.0, generator=g, prompt_embeds=ehs)
    assert isinstance(out, list) and isinstance(out[0], Image.Image) and out[0].size == (32, 32)
    g.manual_seed(7)
    lat = pipe.mask2image_3mod_albedo(" ", np.array([0.3, 0.8]), *imgs, height=32, width=32, num_inference_steps=2,
                                      guidance_scale=0.0, generator=g, prompt_embeds=ehs, output_type="latent")
    assert lat.shape == (1, 4, 8, 8)
    with pytest.raises(NotImplementedError):
        UP.UniRendererPipeline(vae=pipe.vae, unet=pipe.unet, controlnet=pipe.controlnet, controldec=pipe.controldec,
                               safety_checker=object())


def test_assigned_scheduler_config_drives_the_timestep_table(monkeypatch):
    """The eval's `UniPCMultistepScheduler.from_config(pipeline.scheduler.config)` inherits "leading" spacing and
    steps_offset 1 from the SD-1.x base config: the fused loop must walk 941, 894, ... -- and an unsupported option must
    raise instead of being ignored."""
    from uni_renderer_b200 import scheduler as S
    from tests.test_host_logic import SD1X_SCHEDULER_CONFIG
    pipe = _build(monkeypatch)
    base = S.PNDMScheduler.from_config(SD1X_SCHEDULER_CONFIG)
    pipe.scheduler_img = S.UniPCMultistepScheduler.from_config(base.config)
    ehs = torch.randn(1, 7, 48, generator=torch.Generator().manual_seed(3)).half()
    imgs = [_pil(10 + i) for i in range(6)]
    pipe.mask2image_3mod_albedo(" ", np.array([0.3, 0.8]), *imgs, height=32, width=32, num_inference_steps=4,
                                guidance_scale=0.0, generator=torch.Generator().manual_seed(7), prompt_embeds=ehs,
                                output_type="latent")
    plans = [p for p in pipe._sampler._plans.values() if p.scheduler == "unipc"]
    assert len(plans) == 1 and plans[0].timesteps == S.UniPCSchedule.from_config(pipe.scheduler_img.config).timesteps(4)
    assert plans[0].timesteps[0] == 801            # leading: 200 * 4 + 1   (linspace would start at 999)
    # the class-default scheduler (linspace) records a second plan, it does not silently reuse the first
    pipe.scheduler_img = S.UniPCMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    pipe.mask2image_3mod_albedo(" ", np.array([0.3, 0.8]), *imgs, height=32, width=32, num_inference_steps=4,
                                guidance_scale=0.0, generator=torch.Generator().manual_seed(7), prompt_embeds=ehs,
                                output_type="latent")
    assert sorted(p.timesteps[0] for p in pipe._sampler._plans.values() if p.scheduler == "unipc") == [801, 999]
    pipe.scheduler_img = S.UniPCMultistepScheduler(solver_order=3)
    with pytest.raises(NotImplementedError):
        pipe.mask2image_3mod_albedo(" ", np.array([0.3, 0.8]), *imgs, height=32, width=32, num_inference_steps=4,
                                    guidance_scale=0.0, prompt_embeds=ehs)
