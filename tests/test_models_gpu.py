"""Model-level parity on the GPU: the three drop-in modules, called in the reference's 3-call sequence
(train/train.py:1324-1354), against golden tensors recorded from the REFERENCE's own model files
(tests/golden/*.pt, oracle/make_golden.py).

Tolerance.  north_star asks for rtol=1e-3 / atol=1e-4 "fp16"; with fp16 activation storage every layer rounds to
2^-11 relative, so through ~60 sequential layers no fp16 pipeline (the reference's own autocast path included) meets
that elementwise against an fp32 run.  The gate here is therefore: rel_l2 <= 3e-3 on every one of the 66 returned
tensors (measured 0.5-1.3e-3), max abs error <= 1% of the tensor's max magnitude, and -- the yardstick -- our error must
not exceed 1.5x the error of torch's own fp16 execution of the oracle on the same GPU.  Per-op tests
(test_ops_gpu.py) hold the tight per-kernel tolerance."""
import pytest

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("name", ["tiny_step_vec_t.pt", "tiny_step_scalar_t.pt"])
def test_three_call_step_matches_reference_golden(name):
    from tests import gpu_model_probe
    res = gpu_model_probe.run_golden(name)
    assert res.pop("rerun_bit_exact") is True
    assert len(res) == 5 + 36 + 13 + 1
    for k, r in res.items():
        assert r["rel_l2"] <= 3e-3, (k, r)
        assert r["max_abs"] <= 1e-2 * max(r["ref_absmax"], 1e-3), (k, r)


@gpu
def test_error_not_worse_than_torch_fp16():
    """Yardstick: the oracle itself executed in fp16 by torch on the same GPU vs the fp32 golden."""
    import os
    from dataclasses import replace
    import torch
    from oracle import uni_oracle as uo
    from tests import gpu_model_probe
    name = "tiny_step_vec_t.pt"
    ours = gpu_model_probe.run_golden(name)
    gold = torch.load(os.path.join(gpu_model_probe.ROOT, "tests", "golden", name), weights_only=False)
    gc = gold["config"]
    base = uo.NetConfig(block_out_channels=tuple(gc["block_out_channels"]), num_heads=gc["num_heads"],
                        cross_attention_dim=gc["cross_attention_dim"], norm_num_groups=gc["norm_num_groups"])
    cfgs = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [{k: v.cuda().half() for k, v in uo.random_state_dict(kk, c, s).items()}
           for kk, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs, gc["seeds"])]
    B = gc["B"]
    t = torch.full((B,), gc["t_img"], device="cuda")
    orig = uo.timestep_sinusoid
    uo.timestep_sinusoid = lambda tt, dim: orig(tt.cpu(), dim).cuda().half()
    try:
        with torch.no_grad():
            img, attr = uo.dual_stream_step(*sds, *cfgs, gold["x_img"].cuda().half(), t, gold["x_attr"].cuda().half(), t,
                                            gold["ehs"].cuda().half())
    finally:
        uo.timestep_sinusoid = orig
    y_img = gpu_model_probe.err(img, gold["unet_sample"])["rel_l2"]
    y_attr = gpu_model_probe.err(attr, gold["dec_sample"])["rel_l2"]
    assert ours["unet_sample"]["rel_l2"] <= 1.5 * y_img + 2e-4, (ours["unet_sample"], y_img)
    assert ours["dec_sample"]["rel_l2"] <= 1.5 * y_attr + 2e-4, (ours["dec_sample"], y_attr)


@gpu
def test_sd15_shape_step_matches_reference_checksums():
    """BASELINE configs[0] shape (B=1, 64x64 latent, SD-1.5 widths): the reference's outputs were recorded as
    checksums + 64 sampled values per output (tests/golden/sd15_step_checksums.pt)."""
    import os
    import torch
    from tests import gpu_model_probe
    gold = torch.load(os.path.join(gpu_model_probe.ROOT, "tests", "golden", "sd15_step_checksums.pt"), weights_only=False)
    gc = dict(block_out_channels=(320, 640, 1280, 1280), num_heads=8, cross_attention_dim=768, norm_num_groups=32,
              seeds=gold["config"]["seeds"])
    (unet, enc, dec), _, _ = gpu_model_probe.build_modules(gc)
    g = torch.Generator().manual_seed(1234)
    B, S, t = gold["config"]["B"], gold["config"]["S"], gold["config"]["t"]
    x_img = torch.randn(B, 4, S, S, generator=g).cuda()
    x_attr = torch.randn(B, 28, S, S, generator=g).cuda()
    ehs = torch.randn(B, 77, 768, generator=g).cuda()
    d, m, raw_a, raw_a_mid = enc(x_img, t, encoder_hidden_states=ehs, controlnet_cond=x_attr, return_dict=False)
    img, raw_u, raw_u_mid, _ = unet(x_img, t, encoder_hidden_states=ehs, down_block_additional_residuals=d,
                                    mid_block_additional_residual=m, return_dict=False)
    attr = dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=t, encoder_hidden_states=ehs,
               down_block_additional_residuals=raw_u, mid_block_additional_residual=raw_u_mid, return_dict=False)
    torch.cuda.synchronize()
    for out, ref in ((img, gold["img_pred"]), (attr, gold["attr_pred"])):
        o = out.float().cpu()
        vals = o.flatten()[ref["idx"]]
        rel = ((vals - ref["vals"]).norm() / ref["vals"].norm()).item()
        assert rel <= 5e-3, rel
        assert abs(o.norm().item() - ref["l2"]) <= 5e-3 * ref["l2"]
        assert abs(o.std().item() - ref["std"]) <= 5e-3 * ref["std"]
    assert abs(raw_u_mid.float().norm().item() - gold["raw_u_mid_l2"]) <= 5e-3 * gold["raw_u_mid_l2"]
    assert abs(raw_a_mid.float().norm().item() - gold["raw_a_mid_l2"]) <= 5e-3 * gold["raw_a_mid_l2"]


@gpu
def test_upres_decoder_blocks_on_gpu():
    """SURVEY row a8 on the GPU: the class-default UpRes up blocks of AttributeDecoderModel (extra residual after every
    decoder layer) against the oracle; tiny widths, two resolutions of attention blocks + the attention-free block."""
    import torch
    from oracle import uni_oracle as uo
    from tests import gpu_model_probe as gp
    from uni_renderer_b200 import models as M
    gc = dict(block_out_channels=uo.TINY.block_out_channels, num_heads=uo.TINY.num_heads,
              cross_attention_dim=uo.TINY.cross_attention_dim, norm_num_groups=uo.TINY.norm_num_groups, seeds=(11, 12, 13))
    (unet, enc, _), sds, cfgs = gp.build_modules(gc)
    dec = M.AttributeDecoderModel(out_channels=28, _init_weights=False, block_out_channels=tuple(gc["block_out_channels"]),
                                  attention_head_dim=gc["num_heads"], cross_attention_dim=gc["cross_attention_dim"],
                                  norm_num_groups=gc["norm_num_groups"])
    assert dec.net_cfg.up_res
    dec.load_state_dict(sds[2])
    dec.to("cuda")
    g = torch.Generator().manual_seed(1234)
    B, S, t = 2, 16, 501
    x_img, x_attr = torch.randn(B, 4, S, S, generator=g).cuda(), torch.randn(B, 28, S, S, generator=g).cuda()
    ehs = torch.randn(B, 77, cfgs[0].cross_attention_dim, generator=g).cuda()
    d, m, raw_a, raw_a_mid = enc(x_img, t, encoder_hidden_states=ehs, controlnet_cond=x_attr, return_dict=False)
    _, raw_u, raw_u_mid, taps = unet(x_img, t, encoder_hidden_states=ehs, down_block_additional_residuals=d,
                                     mid_block_additional_residual=m, return_dict=False)
    ups = [0.5 * torch.randn(tp.shape, generator=g).cuda() for tp in taps[1:]]
    got = dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=t, encoder_hidden_states=ehs,
              down_block_additional_residuals=raw_u, up_block_additional_residuals=ups,
              mid_block_additional_residual=raw_u_mid, return_dict=False)
    torch.cuda.synchronize()
    c = lambda x: x.float().cpu()          # noqa: E731
    with torch.no_grad():
        ref = uo.attr_decoder_forward(sds[2], cfgs[2], c(raw_a_mid), [c(x) for x in raw_a], t, c(ehs).half().float(),
                                      [c(x) for x in raw_u], c(raw_u_mid),
                                      up_block_additional_residuals=[c(u).half().float() for u in ups])
    r = gp.err(got, ref)
    assert r["rel_l2"] <= 3e-3, r
