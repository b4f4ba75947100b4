"""Classifier-free-guidance plans of the fused loops on the GPU against tests/cfg_reference.py (the reference's guided
step restated on the oracle).  Same teacher-forced per-step gate as tests/test_sampler_gpu.py.  (The file name sorts
last on purpose: this path was added after the round's last GPU slot and is checked on CPU through the op emulator,
tests/test_loops_cpu.py; a surprise here must not hide the results of the other GPU tests under `-x`.)"""
import pytest

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("mode,scheduler,steps", [("joint", "ddim", 1), ("forward", "ddim", 2), ("inverse", "unipc", 3)])
def test_cfg_plan_matches_reference_rules(mode, scheduler, steps):
    import torch
    from oracle import uni_oracle as uo
    from tests import sampler_probe
    from tests.cfg_reference import cfg_step
    sampler, sds, cfgs = sampler_probe.tiny_setup()
    B, S, total, g = 2, 16, 20, 3.0
    gen = torch.Generator().manual_seed(1234)
    x_img, x_attr = torch.randn(B, 4, S, S, generator=gen), torch.randn(B, 28, S, S, generator=gen)
    ehs = torch.randn(B, 77, cfgs[0].cross_attention_dim, generator=gen).half()
    neg = torch.randn(1, 77, cfgs[0].cross_attention_dim, generator=gen).half()
    plan = sampler.plan(mode, B, S, 77, total, scheduler, cfg=True)
    sampler.load_inputs(plan, x_img, x_attr, ehs, neg, g)
    sampler.run(plan, steps=steps)
    torch.cuda.synchronize()
    got_i, got_a = plan.bufs["lat_img"].cpu(), plan.bufs["lat_attr"].cpu()
    mk = (lambda: uo.DDIM()) if scheduler == "ddim" else (lambda: uo.UniPC())
    sched, sched_a = mk(), mk()
    ts = sched.set_timesteps(total)
    sched_a.set_timesteps(total)
    ri, ra = x_img, x_attr
    for i in range(steps):
        ri, ra = cfg_step(mode, sds, cfgs, sched, ts[i], ri, ra, ehs.float(), neg.float(), g, sched_a)
    assert torch.equal(got_a[:, :4], x_attr[:, :4]) and int(plan.bufs["step"].item()) == steps
    if mode != "inverse":
        assert sampler_probe.err(got_i, ri)["rel_l2"] <= 5e-3
    if mode != "forward":
        assert sampler_probe.err(got_a, ra)["rel_l2"] <= 5e-3
