"""The REAL recorders of the fused sampling loops (engine.StreamNet, pipeline.DualStreamSampler.plan: hoisting, lanes,
exchange epilogues, time-embedding tables, fused scheduler updates) executed on the CPU through
tests/cpu_ops_emulator.py and compared with the oracle's step-by-step loop -- host-side wiring of every mode without a
GPU -- and the sharded N > 1 path as a two-process gloo run.  Kernel parity: -m gpu tests."""
import os
import subprocess
import sys
from dataclasses import replace

import pytest
import torch

from oracle import uni_oracle as uo
from tests import cpu_ops_emulator as emu
from tests.sampler_probe import oracle_step

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(prediction_type="epsilon"):
    from uni_renderer_b200.engine import NetConfig
    base = uo.TINY
    cfgs_o = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [uo.random_state_dict(k, c, s) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs_o, (11, 12, 13))]
    nb = NetConfig(block_out_channels=base.block_out_channels, num_heads=base.num_heads,
                   cross_attention_dim=base.cross_attention_dim, norm_num_groups=base.norm_num_groups)
    cfgs = (replace(nb), replace(nb, in_channels=28), replace(nb, out_channels=28))
    return emu.cpu_sampler(sds, cfgs, prediction_type), sds, cfgs_o


def _inputs(B, S, D, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 4, S, S, generator=g), torch.randn(B, 28, S, S, generator=g),
            torch.randn(B, 7, D, generator=g).half())


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("mode,scheduler,steps", [("joint", "ddim", 2), ("forward", "ddim", 2), ("inverse", "ddim", 2),
                                                   ("cycle", "ddim", 2), ("joint", "unipc", 3), ("inverse", "unipc", 3)])
def test_recorded_loop_matches_oracle(monkeypatch, mode, scheduler, steps):
    emu.install(monkeypatch)
    sampler, sds, cfgs = _setup()
    B, S, total = 2, 8, 10
    x_img, x_attr, ehs = _inputs(B, S, cfgs[0].cross_attention_dim)
    plan = sampler.plan(mode, B, S, ehs.shape[1], total, scheduler)
    sampler.load_inputs(plan, x_img, x_attr, ehs)
    sampler.run(plan, steps=steps)
    mk = (lambda: uo.DDIM()) if scheduler == "ddim" else (lambda: uo.UniPC())
    sched, sched_a = mk(), mk()
    ts = sched.set_timesteps(total)
    sched_a.set_timesteps(total)
    ri, ra = x_img, x_attr
    for i in range(steps):
        ri, ra = oracle_step(mode, sds, cfgs, sched, ts[i], ri, ra, ehs.float(), sched_a)
    got_i, got_a = plan.bufs["lat_img"], plan.bufs["lat_attr"]
    assert int(plan.bufs["step"].item()) == steps
    assert torch.equal(got_a[:, :4], x_attr[:, :4])            # the clean mask group is never updated
    if mode != "inverse":
        assert _rel(got_i, ri) < 5e-3, _rel(got_i, ri)
    else:
        assert torch.equal(got_i, x_img)                       # the RGB latent is an input of inverse rendering
    if mode != "forward":
        assert _rel(got_a, ra) < 5e-3, _rel(got_a, ra)
    else:
        assert torch.equal(got_a, x_attr)


def test_plans_are_cached_and_modes_are_validated(monkeypatch):
    emu.install(monkeypatch)
    sampler, _, cfgs = _setup()
    p1 = sampler.plan("joint", 1, 8, 7, 4)
    assert sampler.plan("joint", 1, 8, 7, 4) is p1 and sampler.plan("joint", 1, 8, 7, 5) is not p1
    with pytest.raises(ValueError):
        sampler.plan("sideways", 1, 8, 7, 4)
    with pytest.raises(NotImplementedError):
        sampler.plan("cycle", 1, 8, 7, 4, "unipc")
    with pytest.raises(ValueError):
        sampler.load_inputs(p1, torch.zeros(2, 4, 8, 8), torch.zeros(1, 28, 8, 8), torch.zeros(1, 7, 48))
    with pytest.raises(NotImplementedError):
        sampler.joint_sample(torch.zeros(1, 4, 8, 8), torch.zeros(1, 28, 8, 8), torch.zeros(1, 7, 48).half(), 4,
                             guidance_scale=7.5)
    # step-invariant work is hoisted: forward rendering's step program is shorter than the joint one
    pj, pf = sampler.plan("joint", 1, 8, 7, 4), sampler.plan("forward", 1, 8, 7, 4)
    assert pf.step.num_launches < 0.7 * pj.step.num_launches and pf.setup.num_launches > pj.setup.num_launches


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import pytest
from tests import cpu_ops_emulator as emu
from tests.test_loops_cpu import _setup, _inputs
from uni_renderer_b200.pipeline import all_gather_latents, shard_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mp = pytest.MonkeyPatch()
emu.install(mp)
sampler, sds, cfgs = _setup()
GB, S = 4, 8
x_img, x_attr, ehs = _inputs(GB, S, cfgs[0].cross_attention_dim)          # the same global batch on every rank
lo, hi = shard_batch(GB, rank, world)
img, attr = sampler.joint_sample(x_img[lo:hi], x_attr[lo:hi], ehs[lo:hi], num_inference_steps=2)
out = all_gather_latents(torch.cat([img, attr], 1))                       # the one collective of the sharded loop
if rank == 0:
    full_i, full_a = sampler.joint_sample(x_img, x_attr, ehs, num_inference_steps=2)     # unsharded run of the same batch
    ref = torch.cat([full_i, full_a], 1)
    assert out.shape == ref.shape
    # images are independent; the two runs differ only by fp32 summation order inside the emulator's matmuls (the
    # batch changes the BLAS blocking), which flips a few fp16 roundings: measured rel_l2 2.8e-4
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, rel
    assert torch.equal(out[:, 4:8], torch.cat([x_attr[:, :4]], 0))                       # gather order = batch order
dist.barrier()
dist.destroy_process_group()
mp.undo()
print("OK", rank)
"""


def test_sharded_loop_world2_gloo(tmp_path):
    """Two ranks denoise the two halves of a batch and all-gather the final latents; the result equals the
    single-process run of the whole batch (no cross-sample op anywhere in the loop)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("OK") == 2
