#!/usr/bin/env python
"""One optimizer step of the reference's training loop (train/train.py:1324-1427) at full SD-1.5 widths on the B200
kernels: 3-call dual-stream forward, losses, backward through all three networks, clip, AdamW
(uni_renderer_b200/trainer.py).  Prints one JSON line: seconds per step, images/s, peak memory.  Random-init weights and
synthetic latents (no checkpoints / datasets in this environment).  The attention of this path is the materialised form
(no flash backward yet), so this is a correctness-first number, not a tuned one."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--latent", type=int, default=64)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--tiny", action="store_true", help="the oracle's TINY widths (smoke run)")
ap.add_argument("--cycle", action="store_true", help="add the inverse-rendering consistency pass")
ap.add_argument("--checkpoint", action="store_true", help="activation checkpointing per resnet / transformer block")
ap.add_argument("--graph", action="store_true", help="capture forward + backward into one CUDA graph and replay it")
ap.add_argument("--profile", action="store_true",
                help="bracket the LAST step with cudaProfilerStart/Stop (ncu --profile-from-start off ...)")
a = ap.parse_args()

from dataclasses import replace  # noqa: E402

import torch.distributed as dist  # noqa: E402

# data-parallel training: one process per GPU under torchrun (the reference trains under accelerate DDP); every rank
# holds the same weights, draws its own batch, and the flat gradient buffer is averaged over NCCL in fixed-size buckets
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")

from uni_renderer_b200.engine import NetConfig  # noqa: E402
from uni_renderer_b200.models import random_init_state_dict  # noqa: E402
from uni_renderer_b200.trainer import DualStreamTrainer  # noqa: E402

base = NetConfig(block_out_channels=(32, 64, 128, 128), num_heads=4, cross_attention_dim=48, norm_num_groups=8) if a.tiny \
    else NetConfig()
cfgs = {"unet": replace(base), "enc": replace(base, in_channels=28), "dec": replace(base, out_channels=28)}
kinds = {"unet": "unet", "enc": "attr_enc", "dec": "attr_dec"}
t0 = time.time()
nets = {k: random_init_state_dict(kinds[k], cfgs[k], 3 + i, "cuda", dtype=torch.float32)
        for i, k in enumerate(("unet", "enc", "dec"))}
n_params = sum(v.numel() for sd in nets.values() for v in sd.values())
tr = DualStreamTrainer(nets, cfgs, lr=1e-5, loss_scale=1024.0, max_grad_norm=1.0, gradient_checkpointing=a.checkpoint,
                       use_cuda_graph=a.graph)
del nets
init_s = time.time() - t0
B, S = a.batch, a.latent
g = torch.Generator().manual_seed(1000 * rank)
r = lambda *s: torch.randn(*s, generator=g)                                   # noqa: E731
batch = (r(B, 4, S, S), torch.randint(0, 1000, (B,), generator=g).float(), r(B, 28, S, S),
         torch.randint(0, 1000, (B,), generator=g).float(), r(B, 77, base.cross_attention_dim), r(B, 4, S, S), r(B, 24, S, S))
kw = {}
if a.cycle:
    kw["cycle"] = (r(B, 4, S, S), torch.randint(0, 1000, (B,), generator=g).float())
infos = []
for i in range(a.warmup + a.steps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.time()
    prof = a.profile and i == a.warmup + a.steps - 1
    if prof:
        torch.cuda.profiler.start()
    info = tr.step(*batch, **kw)
    torch.cuda.synchronize()
    if prof:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    info["seconds"] = time.time() - t1
    infos.append(info)
timed = infos[a.warmup:]
sec = sum(i["seconds"] for i in timed) / len(timed)
if world > 1:      # the ranks must hold identical parameters after identical averaged updates
    chk = tr.P.flat[::4097].double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    in_sync = all(float(x) == float(lst[0]) for x in lst)
    t = torch.tensor([sec], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t)
else:
    in_sync = True
if rank == 0:
  print(json.dumps({"n_gpus": world, "ranks_in_sync": in_sync, "collectives_per_step": infos[-1]["collectives"],"what": "3-call dual-stream training step (forward + backward + clip + AdamW), fp16 activations / fp32 master weights",
                  "widths": "tiny" if a.tiny else "SD-1.5", "batch": B, "latent": S, "cycle_pass": bool(a.cycle), "gradient_checkpointing": bool(a.checkpoint), "cuda_graph": bool(a.graph),
                  "parameters": n_params, "seconds_per_step": sec, "images_per_s": world * B / sec,
                  "losses": [round(i["loss"], 5) for i in infos], "grad_norms": [round(i["grad_norm"], 4) for i in infos],
                  "skipped": [i["skipped"] for i in infos], "init_seconds": round(init_s, 1),
                  "peak_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "data": "synthetic"}))
if world > 1:
    dist.destroy_process_group()
