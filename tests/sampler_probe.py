"""GPU probe of the fused sampling loops (uni_renderer_b200.pipeline.DualStreamSampler) against the CPU oracle.

Teacher-forced per-step parity: one denoising step of every mode from identical latents / text embeddings, compared
with the oracle's 3-call step + DDIM update (never a 50-step trajectory elementwise: diffusion trajectories amplify
rounding).  A short multi-step run is compared too, at a looser tolerance.  Used by tests/test_sampler_gpu.py and
__graft_entry__.smoke().
"""
import os
import sys
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import uni_oracle as uo  # noqa: E402  (checker only)


def err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    d = (got - ref).abs()
    return {"max_abs": d.max().item(), "rel_l2": (d.norm() / (ref.norm() + 1e-12)).item(),
            "ref_absmax": ref.abs().max().item(), "finite": bool(torch.isfinite(got).all())}


def tiny_setup(seeds=(11, 12, 13), prediction_type="epsilon", use_graph=True, device="cuda"):
    from uni_renderer_b200.engine import NetConfig
    from uni_renderer_b200.pipeline import DualStreamSampler
    base = uo.TINY
    cfgs_o = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [uo.random_state_dict(k, c, s) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs_o, seeds)]
    nb = NetConfig(block_out_channels=base.block_out_channels, num_heads=base.num_heads,
                   cross_attention_dim=base.cross_attention_dim, norm_num_groups=base.norm_num_groups)
    cfgs = (replace(nb), replace(nb, in_channels=28), replace(nb, out_channels=28))
    sampler = DualStreamSampler.from_state_dicts(*sds, *cfgs, device=device, prediction_type=prediction_type,
                                                 use_graph=use_graph)
    return sampler, sds, cfgs_o


def oracle_step(mode, sds, cfgs, sched, t, x_img, x_attr, ehs, sched_attr=None):
    """One denoising step of `mode` exactly as the reference loops execute it (see bench.py cpu_reference_run).
    Multistep schedulers carry history: pass one instance per stream (sched = RGB stream, sched_attr = attributes),
    like the reference's scheduler_img / scheduler_attr (eval/test_real.py:485-493)."""
    sa = sched_attr if sched_attr is not None else sched
    with torch.no_grad():
        if mode == "forward":
            d, m, _, _ = uo.attr_encoder_forward(sds[1], cfgs[1], 0, ehs, x_attr)
            pred = uo.unet_forward(sds[0], cfgs[0], x_img, t, ehs, d, m)[0]
            return sched.step(pred, t, x_img), x_attr
        if mode == "inverse":
            _, attr = uo.dual_stream_step(*sds, *cfgs, x_img, 0, x_attr, t, ehs)
            return x_img, torch.cat([x_attr[:, :4], sa.step(attr[:, 4:], t, x_attr[:, 4:])], 1)
        img, attr = uo.dual_stream_step(*sds, *cfgs, x_img, t, x_attr, t, ehs)
        if mode == "cycle":
            x2 = torch.cat([x_attr[:, :4], attr[:, 4:]], 1)
            d, m, _, _ = uo.attr_encoder_forward(sds[1], cfgs[1], 0, ehs, x2)
            img = uo.unet_forward(sds[0], cfgs[0], x_img, t, ehs, d, m)[0]
        return sched.step(img, t, x_img), torch.cat([x_attr[:, :4], sa.step(attr[:, 4:], t, x_attr[:, 4:])], 1)


def run_mode(mode, B=2, S=16, steps_total=50, n_steps=1, prediction_type="epsilon", use_graph=True, seed=1234,
             scheduler="ddim", start_index=0):
    """start_index: first step of the walk to replay (e.g. 48 of 50 = t 21, a LATE step where c_out dominates the
    update).  Single DDIM steps also report the error of the recovered network PREDICTION, pred = (x_prev - c_x x) /
    c_out ("img_pred" / "attr_pred"): the latent itself hides the prediction error behind |c_out| ~ 0.02-0.16."""
    sampler, sds, cfgs = tiny_setup(prediction_type=prediction_type, use_graph=use_graph)
    g = torch.Generator().manual_seed(seed)
    x_img = torch.randn(B, 4, S, S, generator=g)
    x_attr = torch.randn(B, 28, S, S, generator=g)
    ehs = torch.randn(B, 77, cfgs[0].cross_attention_dim, generator=g)
    plan = sampler.plan(mode, B, S, 77, steps_total, scheduler)
    sampler.load_inputs(plan, x_img, x_attr, ehs.half())
    if start_index == 0:
        sampler.run(plan, steps=n_steps)
    else:                                    # same as run(), from a later point of the schedule (multistep: DDIM only)
        assert scheduler == "ddim"
        plan.bufs["step"].fill_(start_index)
        plan.setup.run()
        for _ in range(n_steps):
            plan.step.launch_graph() if use_graph else plan.step.run()
    torch.cuda.synchronize()
    got_img, got_attr = plan.bufs["lat_img"].cpu(), plan.bufs["lat_attr"].cpu()
    mk = (lambda: uo.DDIM(prediction_type=prediction_type)) if scheduler == "ddim" else \
        (lambda: uo.UniPC(prediction_type=prediction_type))
    sched, sched_a = mk(), mk()
    ts = sched.set_timesteps(steps_total)
    sched_a.set_timesteps(steps_total)
    ri, ra = x_img, x_attr
    ehs_r = ehs.half().float()          # both sides see the fp16-rounded text embeddings
    for i in range(start_index, start_index + n_steps):
        ri, ra = oracle_step(mode, sds, cfgs, sched, ts[i], ri, ra, ehs_r, sched_a)
    res = {"img": err(got_img, ri), "attr": err(got_attr, ra), "launches_per_step": plan.step.num_launches,
           "mask_untouched": bool(torch.equal(got_attr[:, :4], x_attr[:, :4])),
           "step_counter": int(plan.bufs["step"].item()) - start_index, "t": ts[start_index]}
    if n_steps == 1 and scheduler == "ddim":
        c_out, c_x = sched.coefficients(ts[start_index])
        rec = lambda x_prev, x: (x_prev.double() - c_x * x.double()) / c_out      # noqa: E731
        if mode != "inverse":
            res["img_pred"] = err(rec(got_img, x_img), rec(ri, x_img))
        if mode != "forward":
            res["attr_pred"] = err(rec(got_attr[:, 4:], x_attr[:, 4:]), rec(ra[:, 4:], x_attr[:, 4:]))
    return res


if __name__ == "__main__":
    import json
    mode = sys.argv[1] if len(sys.argv) > 1 else "joint"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    print(json.dumps(run_mode(mode, n_steps=n)))
