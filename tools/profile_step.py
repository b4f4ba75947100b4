#!/usr/bin/env python
"""Runs the sampling loop so that exactly ONE denoising step (eager replay of the recorded step program, every kernel
a separate launch) sits between cudaProfilerStart/Stop -- for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --mode joint
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 3 \
      -o gpurun_out/prof_gemm python tools/profile_step.py --mode joint

Numbers printed under a profiler are never bench values."""
import argparse
import os
import sys
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uni_renderer_b200.engine import NetConfig  # noqa: E402
from uni_renderer_b200.models import random_init_state_dict  # noqa: E402
from uni_renderer_b200.pipeline import DualStreamSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="joint")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--latent", type=int, default=64)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = NetConfig(cross_attention_dim=768)
cfgs = (replace(cfg), replace(cfg, in_channels=28), replace(cfg, out_channels=28))
sds = [random_init_state_dict(k, c, s, dev) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs, (11, 12, 13))]
sampler = DualStreamSampler.from_state_dicts(*sds, *cfgs, device=dev, use_graph=False)
plan = sampler.plan(a.mode, a.batch, a.latent, 77, 50)
g = torch.Generator().manual_seed(1234)
sampler.load_inputs(plan, torch.randn(a.batch, 4, a.latent, a.latent, generator=g),
                    torch.randn(a.batch, 28, a.latent, a.latent, generator=g),
                    torch.randn(a.batch, 77, 768, generator=g).half())
sampler.run(plan, steps=a.warm)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    plan.step.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.steps, "denoising step(s):", plan.step.num_launches, "launches each")
