"""CPU tests of the host-side logic: DDIM schedule vs the oracle, batch sharding + the final all-gather over gloo
(world_size 2), the bench's reference leg on a tiny configuration, and loud failure without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ptype", ["epsilon", "sample", "v_prediction"])
@pytest.mark.parametrize("n", [50, 20, 250])
def test_schedule_matches_oracle_ddim(ptype, n):
    from oracle import uni_oracle as uo
    from uni_renderer_b200.scheduler import DDIMSchedule
    s, o = DDIMSchedule(prediction_type=ptype), uo.DDIM(prediction_type=ptype)
    ts, coefs = s.table(n)
    assert ts == o.set_timesteps(n)
    if n == 50:
        assert ts[0] == 981 and ts[-1] == 1
    g = torch.Generator().manual_seed(0)
    x, out = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for t, (c_out, c_x) in list(zip(ts, coefs))[:: max(1, n // 10)]:
        torch.testing.assert_close(c_out * out + c_x * x, o.step(out, t, x), rtol=2e-5, atol=2e-5)


def test_schedule_rejects_bad_arguments():
    from uni_renderer_b200.scheduler import DDIMSchedule
    with pytest.raises(ValueError):
        DDIMSchedule(prediction_type="nope")
    with pytest.raises(ValueError):
        DDIMSchedule().timesteps(0)


def test_shard_batch():
    from uni_renderer_b200.pipeline import shard_batch
    assert [shard_batch(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from uni_renderer_b200.pipeline import all_gather_latents, shard_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
full = torch.arange(4 * 3 * 2 * 2, dtype=torch.float32).reshape(4, 3, 2, 2)
lo, hi = shard_batch(4, rank, world)
out = all_gather_latents(full[lo:hi] * 1.0)
assert torch.equal(out, full), (rank, out)
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_all_gather_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("OK") == 2


_GRAD_WORKER = """
import sys
sys.path.insert(0, {root!r})
import torch
import torch.distributed as dist
from uni_renderer_b200.trainer import allreduce_gradients
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
n = allreduce_gradients(g, bucket_bytes=4 * 300)          # 300-element buckets: 4 collectives, the last one ragged
assert n == 4, n
assert torch.allclose(g, torch.arange(1000, dtype=torch.float32) * 1.5), (rank, g[:4])
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_bucketed_gradient_allreduce_world2_gloo(tmp_path):
    """Data-parallel training (train/train.py runs under accelerate DDP): the flat gradient buffer is averaged over the
    ranks in fixed-size buckets."""
    script = tmp_path / "g.py"
    script.write_text(_GRAD_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("OK") == 2


def test_allreduce_gradients_is_a_no_op_without_a_process_group():
    from uni_renderer_b200.trainer import allreduce_gradients
    g = torch.ones(10)
    assert allreduce_gradients(g) == 0 and torch.equal(g, torch.ones(10))


def test_sampler_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dataclasses import replace
    from oracle import uni_oracle as uo
    from uni_renderer_b200.engine import NetConfig
    from uni_renderer_b200.pipeline import DualStreamSampler
    from uni_renderer_b200.models import UNet2DConditionModel
    m = UNet2DConditionModel(in_channels=4, out_channels=4, block_out_channels=(32, 64, 128, 128), attention_head_dim=4,
                             cross_attention_dim=48, norm_num_groups=8)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 16, 16), 1, torch.zeros(1, 77, 48))


def test_bench_reference_leg_tiny(monkeypatch):
    """The CPU leg of bench.py (oracle port) on the tiny configuration: runs, returns a positive step time."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle import uni_oracle as uo
    monkeypatch.setattr(uo, "SD15", uo.NetConfig(block_out_channels=(32, 64, 128, 128), num_heads=4,
                                                  cross_attention_dim=768, norm_num_groups=8))
    for mode in ("joint", "forward", "inverse", "cycle"):
        sec, times, cores, sample = bench.cpu_reference_run(mode, 16, 50, timed=1, warm=0)
        assert sec > 0 and len(times) == 1 and cores >= 1 and "oracle" in sample


def test_bench_b200_arm_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=240)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


def test_fold_layernorm_matches_torch_cpu():
    """The host-side algebra of the LayerNorm fold (ops.fold_layernorm, include/unib200.h): with the folded weights,
    wsum and bias, rstd * (x @ wf.T - mean * wsum) + bias2 equals layer_norm(x) @ w.T + b -- checked in fp64-free
    plain torch on the CPU (the only difference allowed is the fp16 rounding of the folded weights)."""
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator().manual_seed(3)
    for (M, C, N) in [(17, 64, 48), (5, 320, 96)]:
        x = torch.randn(M, C, generator=g) * 2.0 + 1.5
        w = torch.randn(N, C, generator=g) * C ** -0.5
        b = torch.randn(N, generator=g)
        gamma = 1.0 + 0.3 * torch.randn(C, generator=g)
        beta = 0.2 * torch.randn(C, generator=g)
        wf, wsum, bias2 = ops.fold_layernorm(w, b, gamma, beta)
        assert torch.equal(wf, wf.half().float()), "folded weights must be fp16-representable"
        mean = x.mean(1, keepdim=True)
        rstd = (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
        got = rstd * (x @ wf.t() - mean * wsum[None]) + bias2[None]
        ref = F.layer_norm(x, (C,), gamma, beta, 1e-5) @ w.t() + b
        torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3)
        # without the fp16 rounding the identity is exact to fp32 round-off
        wf_exact = w * gamma[None]
        got2 = rstd * (x @ wf_exact.t() - mean * wf_exact.sum(1)[None]) + bias2[None]
        torch.testing.assert_close(got2, ref, rtol=1e-4, atol=1e-4)


def test_tile_planning_helpers():
    """pick_bn / rowstats_parts agree with the packing rules (host-only library calls, no GPU needed)."""
    from uni_renderer_b200 import ops
    assert ops.pick_bn(320) == 160 and ops.pick_bn(1280) == 160 and ops.pick_bn(960) == 160
    assert ops.pick_bn(2560, ops.EPI_GEGLU) == 256
    assert ops.pick_bn(4) == 32 and ops.pick_bn(28) == 32
    for n in (32, 64, 128, 320, 640, 1280):
        bn = ops.pick_bn(n)
        assert ops.rowstats_parts(n) == 2 * ((n + bn - 1) // bn)      # two epilogue warps share a row of a tile


@pytest.mark.parametrize("ptype", ["epsilon", "sample", "v_prediction"])
@pytest.mark.parametrize("n", [20, 50, 7])
def test_unipc_tables_match_oracle_stepping(ptype, n):
    """The closed-form [steps][10] UniPC coefficient table (product) drives the same trajectory as the oracle's
    tensor-form multistep algorithm with its model-output history (two independent formulations)."""
    from oracle import uni_oracle as uo
    from uni_renderer_b200.scheduler import UniPCSchedule
    sch, orc = UniPCSchedule(prediction_type=ptype), uo.UniPC(prediction_type=ptype)
    ts, rows = sch.table(n)
    assert ts == orc.set_timesteps(n)
    if n == 20:
        assert ts[0] == 999 and ts[1] == 949 and ts[-1] == 50
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 6, 6, generator=g, dtype=torch.float64)
    S, LS, H0, H1 = x.clone(), torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x)
    ref = x.clone()
    for t, c in zip(ts, rows):
        out = torch.randn(2, 4, 6, 6, generator=g, dtype=torch.float64) * 0.5 + 0.1 * ref     # same net output both sides
        ref_next = orc.step(out, t, ref)
        x0 = c[0] * out + c[1] * S
        Sc = c[2] * LS + c[3] * H0 + c[4] * H1 + c[5] * x0 if c[6] else S
        S, LS, H1, H0 = c[7] * Sc + c[8] * x0 + c[9] * H0, Sc, H0, x0
        torch.testing.assert_close(S, ref_next, rtol=1e-9, atol=1e-9)
        ref = S.clone()


def test_unipc_rejects_bad_arguments():
    from uni_renderer_b200.scheduler import UniPCSchedule
    with pytest.raises(ValueError):
        UniPCSchedule(prediction_type="nope")
    with pytest.raises(ValueError):
        UniPCSchedule().timesteps(0)


# ---------------------------------------------------------------------------------------------------------------
# scheduler configuration honoured from the caller's scheduler objects (eval/test_real.py:485-493)
# ---------------------------------------------------------------------------------------------------------------
SD1X_SCHEDULER_CONFIG = dict(_class_name="PNDMScheduler", beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                             num_train_timesteps=1000, set_alpha_to_one=False, skip_prk_steps=True, steps_offset=1,
                             clip_sample=False, trained_betas=None)          # scheduler_config.json of SD-1.x checkpoints


@pytest.mark.parametrize("spacing,offset,solver", [("leading", 1, "bh2"), ("trailing", 0, "bh2"), ("linspace", 0, "bh1")])
def test_unipc_tables_follow_the_scheduler_config(spacing, offset, solver):
    """`UniPCMultistepScheduler.from_config(pipeline.scheduler.config)` inherits the SD-1.x base config: "leading"
    spacing with steps_offset 1 (941, 894, ... for 20 steps -- not the class default linspace 999, 949, ...)."""
    from oracle import uni_oracle as uo
    from uni_renderer_b200.scheduler import UniPCMultistepScheduler, UniPCSchedule
    holder = UniPCMultistepScheduler.from_config(dict(SD1X_SCHEDULER_CONFIG, timestep_spacing=spacing, steps_offset=offset),
                                                 solver_type=solver)
    sch = UniPCSchedule.from_config(holder.config)
    assert (sch.timestep_spacing, sch.steps_offset, sch.solver_type, sch.beta_schedule) == (spacing, offset, solver,
                                                                                             "scaled_linear")
    orc = uo.UniPC(timestep_spacing=spacing, steps_offset=offset, solver_type=solver)
    n = 20
    ts, rows = sch.table(n)
    assert ts == orc.set_timesteps(n)
    if spacing == "leading":
        assert ts[:2] == [941, 894] and ts[-1] == 48
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 5, 5, generator=g, dtype=torch.float64)
    S, LS, H0, H1, ref = x.clone(), torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x), x.clone()
    for t, c in zip(ts, rows):
        out = torch.randn(1, 4, 5, 5, generator=g, dtype=torch.float64) * 0.5 + 0.1 * ref
        ref_next = orc.step(out, t, ref)
        x0 = c[0] * out + c[1] * S
        Sc = c[2] * LS + c[3] * H0 + c[4] * H1 + c[5] * x0 if c[6] else S
        S, LS, H1, H0 = c[7] * Sc + c[8] * x0 + c[9] * H0, Sc, H0, x0
        torch.testing.assert_close(S, ref_next, rtol=1e-9, atol=1e-9)
        ref = S.clone()


def test_timestep_lists_match_numpy_for_every_step_count():
    """The UniPC "linspace" list must round half-way cases like numpy (n = 30 contains 499, not 500)."""
    import numpy as np
    from uni_renderer_b200.scheduler import DDIMSchedule, UniPCSchedule
    u, d = UniPCSchedule(), DDIMSchedule()
    for n in range(1, 999):
        assert u.timesteps(n) == np.linspace(0, 999, n + 1).round()[::-1][:-1].astype(np.int64).tolist()
    assert 499 in u.timesteps(30) and 500 not in u.timesteps(30)
    for n in (1, 7, 20, 50, 1000):
        assert d.timesteps(n) == [(i * (1000 // n)) + 1 for i in range(n)][::-1]
    dl = DDIMSchedule(timestep_spacing="trailing", steps_offset=0)
    assert dl.timesteps(50)[0] == 999 and dl.timesteps(50)[-1] == 19


def test_ddim_from_config_and_unsupported_options():
    from oracle import uni_oracle as uo
    from uni_renderer_b200.scheduler import DDIMSchedule, DDIMScheduler, UniPCSchedule
    holder = DDIMScheduler.from_config(SD1X_SCHEDULER_CONFIG)
    assert holder.config.steps_offset == 1 and holder.config.clip_sample is False and "skip_prk_steps" not in holder.config
    sch = DDIMSchedule.from_config(holder.config)
    assert sch.signature() == DDIMSchedule().signature()             # == the SD-1.x defaults the bench uses
    orc = uo.DDIM()
    assert sch.timesteps(50) == orc.set_timesteps(50)
    for t in (981, 501, 1):
        a, b = sch.coefficients(t, 50), orc.coefficients(t)
        assert abs(a[0] - b[0]) < 1e-6 and abs(a[1] - b[1]) < 1e-6
    tr = DDIMSchedule.from_config(dict(holder.config, timestep_spacing="trailing"))
    otr = uo.DDIM(timestep_spacing="trailing")
    assert tr.timesteps(20) == otr.set_timesteps(20)
    with pytest.raises(NotImplementedError):
        DDIMSchedule.from_config(dict(holder.config, clip_sample=True))        # diffusers' own DDIM default
    with pytest.raises(NotImplementedError):
        UniPCSchedule.from_config(dict(SD1X_SCHEDULER_CONFIG, solver_order=3))
    with pytest.raises(NotImplementedError):
        UniPCSchedule.from_config(dict(SD1X_SCHEDULER_CONFIG, use_karras_sigmas=True))
    with pytest.raises(NotImplementedError):
        DDIMSchedule.from_config(dict(holder.config, beta_schedule="sigmoid"))


def test_pipeline_from_pretrained_loads_the_scheduler_config(tmp_path):
    """`pipeline.scheduler.config` must exist after from_pretrained: the eval builds every per-stream scheduler from it."""
    import json
    from uni_renderer_b200.scheduler import UniPCMultistepScheduler, UniPCSchedule, load_scheduler_config
    (tmp_path / "scheduler").mkdir()
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(SD1X_SCHEDULER_CONFIG))
    base = load_scheduler_config(str(tmp_path / "scheduler"))
    assert type(base).__name__ == "PNDMScheduler" and base.config.steps_offset == 1
    per_stream = UniPCMultistepScheduler.from_config(base.config)
    assert per_stream.config.timestep_spacing == "leading" and per_stream.config.solver_order == 2
    assert UniPCSchedule.from_config(per_stream.config).timesteps(20)[0] == 941
