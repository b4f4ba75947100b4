#!/usr/bin/env python
"""Per-op device-time table of ONE denoising step (CUDA events around every op, unib200_program_profile), grouped by
shape.  Writes JSON to --out.  Optimisation instrument; not a bench value (ops run back to back, eagerly)."""
import argparse
import json
import os
import sys
from collections import OrderedDict
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
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/ops.json")
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
sampler.run(plan, steps=2)
torch.cuda.synchronize()
ms = plan.step.profile(a.iters)
info = plan.step.op_info()
groups = OrderedDict()
for i, ((kind, fl, by, nl), t) in enumerate(zip(info, ms)):
    d = plan.step.op_desc(i)
    gq = groups.setdefault(d, {"n": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
    gq["n"] += 1; gq["ms"] += t; gq["flops"] += fl; gq["bytes"] += by
rows = []
for d, v in groups.items():
    rows.append({"op": d, "count": v["n"], "ms_total": round(v["ms"], 4), "us_each": round(1e3 * v["ms"] / v["n"], 2),
                 "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                 "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None})
rows.sort(key=lambda r: -r["ms_total"])
tot = sum(r["ms_total"] for r in rows)
os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
with open(a.out, "w") as f:
    json.dump({"mode": a.mode, "batch": a.batch, "latent": a.latent, "total_ms": tot, "rows": rows}, f, indent=1)
print(f"total {tot:.3f} ms over {len(ms)} ops")
for r in rows[:60]:
    print(f"{r['ms_total']:8.3f} ms  x{r['count']:<3d} {r['us_each']:8.1f} us  {str(r['tflops']):>7} TF  {str(r['gbs']):>7} GB/s  {r['op']}")
