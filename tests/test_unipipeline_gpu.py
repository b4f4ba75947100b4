"""`UniRendererPipeline` on the GPU, the way eval/test_real.py:470-553 drives the reference pipeline: components passed
as keywords, per-stream UniPC schedulers derived from the base scheduler's config (which makes them "leading" /
steps_offset 1), PIL image + mask in -> material latents + five PIL maps out -- against the oracle chain
(vae_oracle.encode -> uni_oracle UniPC steps with the SAME scheduler configuration -> vae_oracle.decode)."""
from dataclasses import replace

import numpy as np
import pytest

gpu = pytest.mark.gpu
TINY = dict(block_out_channels=(32, 64, 128, 128), attention_head_dim=4, cross_attention_dim=48, norm_num_groups=8)


def _pil(seed, size=72):
    from PIL import Image
    rng = np.random.default_rng(seed)
    return Image.fromarray(rng.integers(0, 256, (size, size, 3), dtype=np.uint8))


@gpu
def test_real_image2mask_pil_roundtrip_matches_oracle_chain():
    import torch
    from PIL import Image
    from oracle import uni_oracle as uo
    from oracle import vae_oracle as vo
    from tests import sampler_probe
    from tests.test_host_logic import SD1X_SCHEDULER_CONFIG
    from uni_renderer_b200 import models as M
    from uni_renderer_b200 import scheduler as S
    from uni_renderer_b200 import unipipeline as UP
    from uni_renderer_b200 import vae as V
    base = uo.TINY
    cfgs = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [uo.random_state_dict(k, c, s) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs, (11, 12, 13))]
    unet = M.UNet2DConditionModel(in_channels=4, out_channels=4, _init_weights=False, **TINY)
    enc = M.AttributeEncoderModel(in_channels=28, _init_weights=False, **TINY)
    dec = M.AttributeDecoderModel(out_channels=28, up_block_types=M._SD_UP, _init_weights=False, **TINY)
    for m, sd in zip((unet, enc, dec), sds):
        m.load_state_dict(sd)
    vsd = vo.random_state_dict(vo.TINY_VAE, 5)
    vae = V.AutoencoderKL(block_out_channels=(32, 64, 64), down_block_types=(V._DOWN,) * 3, up_block_types=(V._UP,) * 3,
                          layers_per_block=2, norm_num_groups=8)
    vae.load_state_dict(vsd)
    pipe = UP.UniRendererPipeline.from_pretrained(None, vae=vae, text_encoder=None, tokenizer=None, unet=unet,
                                                  controlnet=enc, controldec=dec, safety_checker=None,
                                                  scheduler=S.PNDMScheduler.from_config(SD1X_SCHEDULER_CONFIG))
    pipe = pipe.to("cuda")
    for n in UP._STREAM_SCHEDULERS:                      # eval/test_real.py:485-493
        setattr(pipe, n, S.UniPCMultistepScheduler.from_config(pipe.scheduler.config))
    pipe.set_progress_bar_config(disable=True)
    H, steps, h = 64, 3, 16
    ehs = torch.randn(1, 77, 48, generator=torch.Generator().manual_seed(3)).half()
    gen = torch.Generator(device="cuda").manual_seed(42)
    torch.cuda.manual_seed(9)            # the posterior noise of latent_dist.sample() comes from the global RNG
    material, normal, albedo, spec, diff, env = pipe.real_image2mask_3mod_albedo(
        " ", _pil(1), _pil(2), guidance_scale=0.0, height=H, width=H, num_inference_steps=steps, generator=gen,
        prompt_embeds=ehs)
    torch.cuda.synchronize()
    assert material.shape == (1, 4, h, h) and torch.isfinite(material).all()
    for out in (normal, albedo, spec, diff, env):
        assert isinstance(out, list) and isinstance(out[0], Image.Image) and out[0].size == (H, H)
    plan = [p for p in pipe._sampler._plans.values() if p.mode == "inverse"][0]
    assert plan.scheduler == "unipc" and plan.timesteps == [751, 501, 251]      # leading, offset 1: not 999, 666, 333

    # ---- oracle chain with the same draws
    x, msk = UP.preprocess_image(_pil(1), H, H), UP.preprocess_image(_pil(2), H, H)
    torch.cuda.manual_seed(9)
    n_img, n_msk = (torch.randn(1, 4, h, h, device="cuda").cpu() for _ in range(2))
    gen.manual_seed(42)
    lat = [torch.randn(1, 4, h, h, generator=gen, device="cuda").cpu() for _ in range(6)]
    sf = vo.TINY_VAE.scaling_factor
    with torch.no_grad():
        l_img = vo.sample_posterior(vo.encode_moments(vsd, vo.TINY_VAE, x), n_img) * sf
        l_msk = vo.sample_posterior(vo.encode_moments(vsd, vo.TINY_VAE, msk), n_msk) * sf
        mk = lambda: uo.UniPC(timestep_spacing="leading", steps_offset=1)      # noqa: E731
        sched, sched_a = mk(), mk()
        ts = sched.set_timesteps(steps)
        sched_a.set_timesteps(steps)
        assert ts == plan.timesteps
        xi, xa = l_img, torch.cat([l_msk] + lat, 1)
        for i in range(steps):
            xi, xa = sampler_probe.oracle_step("inverse", sds, cfgs, sched, ts[i], xi, xa, ehs.float(), sched_a)
        rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()    # noqa: E731
        assert rel(material, xa[:, 4:8]) <= 5e-3, rel(material, xa[:, 4:8])
        for i, out in enumerate((normal, albedo, spec, diff, env)):
            ref = vo.decode(vsd, vo.TINY_VAE, xa[:, 8 + 4 * i:12 + 4 * i] / sf)
            ref_u8 = np.asarray(UP.postprocess_image(ref, "pil")[0]).astype(np.int32)
            got_u8 = np.asarray(out[0]).astype(np.int32)
            assert np.abs(got_u8 - ref_u8).max() <= 3 and np.abs(got_u8 - ref_u8).mean() <= 0.6, (i, np.abs(got_u8 - ref_u8).max())
