#!/usr/bin/env python
"""Benchmark of the dual-stream denoising hot path (BASELINE.json: 512x512 dual-stream 50-step DDIM images/sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's arithmetic on the host cores (CPU)

A bench "step" = one full 50-step DDIM sampling pass of one batch (default: BASELINE configs[1], B=4 at 64x64 latent,
joint dual-stream, fp16 storage / fp32 accumulate, random-init SD-1.5-shaped networks, synthetic latents).
  value   images/sec with the inputs already resident in HBM (device-to-device reset of the latents per step)
  e2e     images/sec through DualStreamSampler.joint_sample()-style calls with PINNED HOST inputs: H2D of latents and
          text embeddings and D2H of the final latents inside the timed region (+ the NCCL all-gather for N > 1)
  roofline  the dominant kernel family (tcgen05 implicit-GEMM conv/linear): algorithmic FLOPs of its launches in one
          denoising step / their summed device time (CUDA events around every launch, measured here), against the
          measured sustained tensor peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/uni_oracle.py, the restatement of the reference's PyTorch path) timed on this
          box's host cores on a bounded sample (B=1, a few denoising steps, extrapolated to 50)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec, 512x512 dual-stream 50-step DDIM"
UNIT = "images/s"
MODE_CONFIG = {"joint": "configs[1]: batch=4 512x512 dual-stream 50-step DDIM, fp16, 1xB200",
               "forward": "configs[2] per-GPU shard: forward rendering (attribute->RGB) 50-step, batch 4/GPU",
               "inverse": "configs[3] per-GPU shard: inverse rendering (RGB->attributes) 50-step, batch 4/GPU",
               "cycle": "configs[4] per-GPU shard: 1024x1024 dual-stream + cycle-consistency double pass, batch 2/GPU"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="joint", choices=["joint", "forward", "inverse", "cycle"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 4; 2 for --mode cycle)")
    ap.add_argument("--latent", type=int, default=None, help="latent side (default 64; 128 for --mode cycle)")
    ap.add_argument("--denoise-steps", type=int, default=50)
    ap.add_argument("--scheduler", default="ddim", choices=["ddim", "unipc"],
                    help="ddim = BASELINE.json's metric; unipc = the scheduler of the reference's shipped eval")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--split-batch", action="store_true",
                    help="forward/inverse: the two batch halves on two lanes (measured slower at B=4; off by default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-vae", action="store_true", help="skip the AutoencoderKL side measurement (N=1, 64x64 latents)")
    ap.add_argument("--no-modes", action="store_true",
                    help="skip the `modes` legs (forward / inverse / cycle shards of BASELINE configs[2..4])")
    ap.add_argument("--no-train", action="store_true",
                    help="skip the training-step side measurement (tools/bench_train.py as a subprocess, N=1)")
    ap.add_argument("--no-torch-eager", action="store_true",
                    help="skip the live torch-eager fp16 GPU baseline leg (oracle/torch_eager.py as a subprocess)")
    ap.add_argument("--cpu-denoise-steps", type=int, default=2, help="timed CPU denoising steps of the cpu_baseline leg")
    a = ap.parse_args()
    if a.warmup < 3:
        a.warmup = 3
    if a.batch is None:
        a.batch = 2 if a.mode == "cycle" else 4
    if a.latent is None:
        a.latent = 128 if a.mode == "cycle" else 64
    return a


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"tflops_sustained": d.get("bf16_tflops_sustained", 1402.9), "tflops_burst": d.get("bf16_tflops", 1664.2),
                "hbm_gbs": d.get("hbm_gbs", 6551.7), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# --------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle (restatement of the reference's PyTorch path) on the host cores
# --------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(mode: str, latent: int, denoise_steps: int, timed: int, warm: int):
    """Times `timed` denoising steps (after `warm`) of the reference arithmetic at B=1 on all host cores; returns
    (seconds per denoising step, cores, description).  Uses oracle/ -- the one place bench.py may do so."""
    import torch
    from dataclasses import replace
    from oracle import uni_oracle as uo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    base = uo.SD15
    cfgs = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [uo.random_state_dict(k, c, s) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs, (11, 12, 13))]
    g = torch.Generator().manual_seed(1234)
    x_img = torch.randn(1, 4, latent, latent, generator=g)
    x_attr = torch.randn(1, 28, latent, latent, generator=g)
    ehs = torch.randn(1, 77, 768, generator=g)
    sched = uo.DDIM()
    ts = sched.set_timesteps(denoise_steps)

    def one(i, x_img, x_attr):
        t = ts[i]
        if mode == "forward":       # pipeline.py:1586-1653 as shipped (the attribute encoder is re-run every step)
            d, m, _, _ = uo.attr_encoder_forward(sds[1], cfgs[1], 0, ehs, x_attr)
            pred = uo.unet_forward(sds[0], cfgs[0], x_img, t, ehs, d, m)[0]
            return sched.step(pred, t, x_img), x_attr
        if mode == "inverse":       # pipeline.py:2627-2733 as shipped (full 3-call step, RGB output discarded)
            _, attr = uo.dual_stream_step(*sds, *cfgs, x_img, 0, x_attr, t, ehs)
            x_attr = torch.cat([x_attr[:, :4], sched.step(attr[:, 4:], t, x_attr[:, 4:])], 1)
            return x_img, x_attr
        img, attr = uo.dual_stream_step(*sds, *cfgs, x_img, t, x_attr, t, ehs)
        if mode == "cycle":         # train.py:1388-1413
            x2 = torch.cat([x_attr[:, :4], attr[:, 4:]], 1)
            d, m, _, _ = uo.attr_encoder_forward(sds[1], cfgs[1], 0, ehs, x2)
            img = uo.unet_forward(sds[0], cfgs[0], x_img, t, ehs, d, m)[0]
        x_attr = torch.cat([x_attr[:, :4], sched.step(attr[:, 4:], t, x_attr[:, 4:])], 1)
        return sched.step(img, t, x_img), x_attr

    times = []
    with torch.no_grad():
        for i in range(warm + timed):
            t0 = time.perf_counter()
            x_img, x_attr = one(i % denoise_steps, x_img, x_attr)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    sec = sum(times) / len(times)
    sample = (f"oracle fp32 PyTorch-CPU, B=1, {latent}x{latent} latent, mode={mode}: {warm} warm + {timed} timed "
              f"denoising steps, extrapolated x{denoise_steps} steps per image")
    return sec, times, cores, sample


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, times, cores, sample = cpu_reference_run(a.mode, a.latent, a.denoise_steps, a.steps, a.warmup)
    value = 1.0 / (sec * a.denoise_steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": MODE_CONFIG[a.mode], "mode": a.mode, "batch": 1, "latent": a.latent,
                       "denoise_steps": a.denoise_steps,
                       "note": "one bench step of this arm = ONE denoising step at B=1 (bounded sample); value = "
                               "1 / (50 x mean step seconds)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from uni_renderer_b200 import _lib
    from uni_renderer_b200.engine import NetConfig
    from uni_renderer_b200.models import random_init_state_dict
    from uni_renderer_b200.pipeline import DualStreamSampler, all_gather_latents

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU leg)")
    _lib.load()                                  # fail loudly when the extension is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, T, L = a.batch, a.latent, a.denoise_steps, 77

    cfg = NetConfig(cross_attention_dim=768)
    from dataclasses import replace
    cfgs = (replace(cfg), replace(cfg, in_channels=28), replace(cfg, out_channels=28))
    sds = [random_init_state_dict(k, c, s, dev) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs, (11, 12, 13))]
    sampler = DualStreamSampler.from_state_dicts(*sds, *cfgs, device=dev, use_graph=not a.no_graph,
                                                 split_batch=a.split_batch)
    del sds
    plan = sampler.plan(a.mode, B, S, L, T, a.scheduler)
    torch.cuda.synchronize()

    # synthetic inputs: per-rank seeds so every rank denoises different images (weak scaling: B per GPU is fixed)
    g = torch.Generator().manual_seed(1234 + rank)
    h_img = torch.randn(B, 4, S, S, generator=g).pin_memory()
    h_attr = torch.randn(B, 28, S, S, generator=g).pin_memory()
    h_ehs = torch.randn(B, L, 768, generator=g).half().pin_memory()
    d_img, d_attr, d_ehs = h_img.to(dev), h_attr.to(dev), h_ehs.to(dev)
    out_c = {"joint": 28 + 4, "cycle": 28 + 4, "forward": 4, "inverse": 24}[a.mode]
    h_out = torch.empty(world * B, out_c, S, S).pin_memory() if rank == 0 else None

    def final_latents():
        b = plan.bufs
        if a.mode == "forward":
            return b["lat_img"]
        if a.mode == "inverse":
            return b["lat_attr"][:, 4:].contiguous()
        return torch.cat([b["lat_img"], b["lat_attr"]], 1)

    def step_resident():
        sampler.load_inputs(plan, d_img, d_attr, d_ehs)
        sampler.run(plan)

    def step_e2e():
        sampler.load_inputs(plan, h_img, h_attr, h_ehs)            # H2D from pinned memory
        sampler.run(plan)
        out = all_gather_latents(final_latents())                  # the one collective (no-op at N=1)
        if rank == 0:
            h_out.copy_(out, non_blocking=True)                    # D2H of the final latents

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(a.warmup):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    ms_res = timed(step_resident, a.steps)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    clk = clocks.stop()

    images = world * B * a.steps
    value = images / (ms_res * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    h2d = h_img.numel() * 4 + h_attr.numel() * 4 + h_ehs.numel() * 2
    d2h = world * B * out_c * S * S * 4

    peaks = load_peaks()
    flops_call = plan.flops_setup + T * plan.flops_step
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": MODE_CONFIG[a.mode] + (f" (weak scaling: the same shard on each of {world} GPUs)"
                                                          if world > 1 else ""),
                       "mode": a.mode, "per_gpu_batch": B, "global_batch": world * B,
                       "latent": S, "denoise_steps": T, "scheduler": a.scheduler, "text_tokens": L, "weights": "random-init SD-1.5 shape "
                       "(859.5M + 360.3M + 524.4M params)", "cuda_graph": not a.no_graph,
                       "l2": "every denoising step streams 3.5 GB of weights + activations >> 126 MB L2; no flush needed"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": a.steps * sampler.launches_per_call(plan),
            "denoise_step_ms": ms_res / a.steps / T,
            "step_tflops_per_gpu": flops_call * a.steps / (ms_res * 1e-3) / 1e12,
            "step_frac_of_tensor_peak": flops_call * a.steps / (ms_res * 1e-3) / 1e12 / peaks["tflops_sustained"]}

    if rank == 0 and a.mode == "joint" and not a.no_roofline and T > 5:
        # BASELINE.json's second figure, "UNet ms/step": the RGB UNet alone (the step of the forward-rendering loop:
        # UNet encoder + exchange adds + decoder + DDIM update, attribute encoder hoisted) at the same batch, CUDA events
        uplan = sampler.plan("forward", B, S, L, T)
        sampler.load_inputs(uplan, d_img, d_attr, d_ehs)
        sampler.run(uplan, steps=3)                    # setup program (hoisted attribute encoder) + warm-up steps
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(T - 3):                         # the remaining steps of the same 50-step walk
            uplan.step.launch_graph() if sampler.use_graph else uplan.step.run()
        u1.record()
        torch.cuda.synchronize()
        line["unet_ms_per_step"] = u0.elapsed_time(u1) / (T - 3)
        line["unet_ms_per_step_note"] = (f"RGB UNet forward (encoder + exchange adds + decoder) + fused DDIM update, B={B}, "
                                         f"{S}x{S} latent: mean of {T - 3} step-graph replays, CUDA events")
    if rank == 0 and not a.no_roofline:
        # per-op device times of ONE denoising step (events around every launch on the launching stream)
        sampler.load_inputs(plan, d_img, d_attr, d_ehs)
        plan.bufs["step"].zero_()
        plan.setup.run()
        ms_ops = plan.step.profile(3)
        info = plan.step.op_info()
        by = {}
        for (kind, fl, by_, nl), ms in zip(info, ms_ops):
            d = by.setdefault(_lib.OP_NAMES[kind], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += ms; d["flops"] += fl; d["bytes"] += by_; d["launches"] += nl
        tot = sum(d["ms"] for d in by.values())
        gm = by["conv_gemm"]
        ach = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
        # DRAM traffic of the same kernel family from the committed ncu capture of one denoising step (profiles/):
        # per launch, like `achieved`; ncu replays each launch cold-cache, so L2-resident activations count as DRAM
        traffic, traffic_src = None, None
        import glob
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_traffic.json")))
        if a.mode == "joint" and B == 4 and S == 64 and cands:
            with open(cands[-1]) as f:           # newest capture (files are named per round: r1b_, r1c_, ...)
                tj = json.load(f)
            traffic = tj["dram_bytes_per_launch"]
            traffic_src = f"profiles/{os.path.basename(cands[-1])} (ncu, cold-cache replay)"
        line["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (implicit-GEMM conv / linear family)",
                            "achieved": ach, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                            "frac": ach / peaks["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                            "algorithmic_bytes_per_launch": gm["bytes"] / max(gm["launches"], 1),
                            "algorithmic_flops_per_launch": gm["flops"] / max(gm["launches"], 1),
                            "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                            "launches_per_denoise_step": gm["launches"], "share_of_step": gm["ms"] / tot,
                            "flops_per_denoise_step": gm["flops"]}
        line["kernel_breakdown_ms_per_denoise_step"] = {
            k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 4), "launches": v["launches"],
                "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None}
            for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}
    # BASELINE configs[2..4]: the per-GPU shards of forward rendering, inverse rendering (global batch 32 on 8 GPUs = 4 per
    # GPU) and the 1024x1024 cycle double pass (global 16 = 2 per GPU), measured at THIS N through the same end-to-end
    # path as `e2e` (pinned host inputs, H2D, the loop, the one all-gather, D2H) -- so the driver's N = 1..8 scaling
    # runs record them too.  Reported beside the headline, never inside it.
    if a.mode == "joint" and not a.no_modes and a.scheduler == "ddim" and not a.no_graph:
        line["modes"] = {}
        for m2, (B2, S2) in (("forward", (4, 64)), ("inverse", (4, 64)), ("cycle", (2, 128))):
            try:
                line["modes"][m2] = mode_leg(torch, dist, sampler, m2, B2, S2, L, T, dev, rank, world, all_gather_latents)
            except Exception as e:  # noqa: BLE001  (a side leg must not take the headline down)
                line["modes"][m2] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
    if rank == 0 and a.mode == "joint" and B == 4 and S == 64 and a.scheduler == "ddim" and not a.no_torch_eager:
        # SURVEY 8d's "same-box GPU PyTorch" bar, measured LIVE on this box: the reference's arithmetic under torch
        # eager fp16 (cuDNN / cuBLAS / SDPA), as a separate process outside every timed region of this arm
        line["gpu_pytorch_eager"] = torch_eager_leg(local, B, S, ms_res / a.steps / T)
    if rank == 0 and world == 1 and a.mode == "joint" and B == 4 and S == 64 and not a.no_train:
        # one optimizer step of the reference's training loop (3-call forward + backward + clip + AdamW,
        # train/train.py:1324-1427) at full widths on the same kernels; a separate process, beside the headline
        line["train"] = train_leg(local)
    if rank == 0 and world == 1 and not a.no_vae and S == 64:
        # the next row of the scope table: the AutoencoderKL that brackets every sampling call of the
        # reference (models/pipeline.py:1531-1556 encodes, :1664 / :2335-2349 decodes) on the same kernels.  Reported
        # beside the headline, never inside it: BASELINE.json's metric is the denoising loop on latents.
        try:
            line["vae"] = vae_leg(torch, dev, B, S * 8, peaks)
        except Exception as e:  # noqa: BLE001  (the headline must survive a failure of the side measurement)
            line["vae"] = {"error": f"{type(e).__name__}: {e}"}
    barrier()

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sec, _, cores, sample = cpu_reference_run(a.mode, S, T, a.cpu_denoise_steps, 1)
        line["cpu_baseline"] = {"value": 1.0 / (sec * T), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "sec_per_denoise_step_b1": sec}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def mode_leg(torch, dist, sampler, mode, B, S, L, T, dev, rank, world, all_gather_latents, passes: int = 2):
    """One of BASELINE configs[2..4] at this N: 1 warm + `passes` timed end-to-end sampling passes of the per-GPU shard
    (pinned host inputs -> H2D -> setup + T denoising steps -> all-gather of the final latents -> D2H on rank 0)."""
    plan = sampler.plan(mode, B, S, L, T)
    g = torch.Generator().manual_seed(4321 + rank)
    h_img = torch.randn(B, 4, S, S, generator=g).pin_memory()
    h_attr = torch.randn(B, 28, S, S, generator=g).pin_memory()
    h_ehs = torch.randn(B, L, 768, generator=g).half().pin_memory()
    out_c = {"cycle": 32, "forward": 4, "inverse": 24}[mode]
    h_out = torch.empty(world * B, out_c, S, S).pin_memory() if rank == 0 else None

    def one():
        sampler.load_inputs(plan, h_img, h_attr, h_ehs)
        sampler.run(plan)
        b = plan.bufs
        fin = b["lat_img"] if mode == "forward" else (b["lat_attr"][:, 4:].contiguous() if mode == "inverse" else
                                                      torch.cat([b["lat_img"], b["lat_attr"]], 1))
        out = all_gather_latents(fin)
        if rank == 0:
            h_out.copy_(out, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        one()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    finite = bool(torch.isfinite(plan.bufs["lat_img"]).all() and torch.isfinite(plan.bufs["lat_attr"]).all())
    res = {"workload": MODE_CONFIG[mode], "value": world * B * passes / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
           "per_gpu_batch": B, "global_batch": world * B, "latent": S, "denoise_steps": T, "passes": passes,
           "ms_per_pass": ms / passes, "denoise_step_ms": ms / passes / T, "timing": "end to end (H2D, loop, all-gather, D2H)",
           "launches_per_pass": sampler.launches_per_call(plan), "finite": finite}
    # the plan's buffers (activations of a second shape) are released: the headline plan stays resident
    for k in [k for k, p_ in sampler._plans.items() if p_ is plan]:
        del sampler._plans[k]
    return res


def torch_eager_leg(gpu_index: int, B: int, S: int, our_ms_per_denoise_step: float):
    """oracle/torch_eager.py as a subprocess on the same GPU: (a) cuDNN heuristics, NCHW -- how the reference's eval runs;
    (b) cudnn.benchmark autotuning + channels_last -- the best a torch-eager user gets.  Each bounded by a timeout."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[gpu_index]
               if os.environ.get("CUDA_VISIBLE_DEVICES") else str(gpu_index))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    out = {"what": "torch eager fp16 (cuDNN / cuBLAS / SDPA) joint dual-stream denoising step, same box, measured live",
           "batch": B, "latent": S, "runs": []}
    for flags, tmo in (([], 150), (["--cudnn-benchmark", "--channels-last"], 240)):
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "torch_eager.py"), "--batch", str(B), "--latent", str(S),
               "--steps", "5", "--warmup", "3"] + flags
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=tmo, env=env)
            ln = [x for x in p.stdout.splitlines() if x.startswith("{")]
            if p.returncode == 0 and ln:
                d = json.loads(ln[-1])
                out["runs"].append({"cudnn_benchmark": d["cudnn_benchmark"], "channels_last": d.get("channels_last", False),
                                    "ms_per_denoise_step": d["ms_per_denoise_step"], "torch": d["torch"],
                                    "finite": d["finite"]})
            else:
                out["runs"].append({"flags": flags, "error": (p.stderr or p.stdout)[-300:]})
        except subprocess.TimeoutExpired:
            out["runs"].append({"flags": flags, "error": f"timeout after {tmo} s"})
    ok = [r for r in out["runs"] if "ms_per_denoise_step" in r]
    if ok:
        best = min(ok, key=lambda r: r["ms_per_denoise_step"])
        out["ms_per_denoise_step"] = best["ms_per_denoise_step"]
        out["images_per_s"] = B / (50 * best["ms_per_denoise_step"] * 1e-3)
        out["speedup_vs_it"] = best["ms_per_denoise_step"] / our_ms_per_denoise_step
    return out


def train_leg(gpu_index: int):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[gpu_index]
               if os.environ.get("CUDA_VISIBLE_DEVICES") else str(gpu_index))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "bench_train.py"), "--batch", "2", "--latent", "64", "--steps", "3",
           "--warmup", "2", "--graph"]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
        ln = [x for x in p.stdout.splitlines() if x.startswith("{")]
        if p.returncode == 0 and ln:
            d = json.loads(ln[-1])
            d["parity"] = "tests/test_trainer_gpu.py (every parameter gradient vs torch autograd of the oracle)"
            return d
        return {"error": (p.stderr or p.stdout)[-300:]}
    except subprocess.TimeoutExpired:
        return {"error": "timeout after 240 s"}


def vae_leg(torch, dev, B, image, peaks):
    """One batched AutoencoderKL decode and encode (SD-1.x widths, random init) at the bench's batch: device ms (CUDA
    events, 3 warm + 5 timed), algorithmic TFLOP/s and the per-class split of an eager replay."""
    from uni_renderer_b200 import _lib
    from uni_renderer_b200 import vae as V
    cfg = V.VaeConfig()
    m = V.AutoencoderKL(block_out_channels=cfg.block_out_channels, down_block_types=(V._DOWN,) * 4,
                        up_block_types=(V._UP,) * 4, layers_per_block=2, norm_num_groups=32, _init_weights=False)
    m.load_state_dict(V.random_init_vae_state_dict(cfg, 21, dev))
    m = m.to(dev)
    g = torch.Generator(device=dev).manual_seed(4)
    z = torch.randn(B, 4, image // 8, image // 8, generator=g, device=dev)
    x = torch.tanh(torch.randn(B, 3, image, image, generator=g, device=dev))
    out = {"config": f"AutoencoderKL SD-1.x widths (83.7M params, random init), batch {B}, {image}x{image} images, fp16 "
                     "storage / fp32 accumulate", "parity": "tests/test_vae_gpu.py (oracle/vae_oracle.py, parity unpinned)"}
    for name, key, fn in (("decode", ("dec", B, image // 8, image // 8), lambda: m.decode(z)),
                          ("encode", ("enc", B, image, image), lambda: m.encode(x))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        P = m._progs[key]["prog"]
        info, ms_ops = P.op_info(), P.profile(3)
        by = {}
        for (kind, fl, by_, nl), t in zip(info, ms_ops):
            d = by.setdefault(_lib.OP_NAMES[kind], {"ms": 0.0, "flops": 0.0, "launches": 0})
            d["ms"] += t; d["flops"] += fl; d["launches"] += nl
        flops = sum(i[1] for i in info)
        out[name] = {"ms": ms, "images_per_s": B / (ms * 1e-3), "launches": P.num_launches,
                     "tflops": flops / (ms * 1e-3) / 1e12,
                     "frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"],
                     "by_kind_ms": {k: round(v["ms"], 3) for k, v in by.items()},
                     "gemm_tflops_eager": round(by["conv_gemm"]["flops"] / (by["conv_gemm"]["ms"] * 1e-3) / 1e12, 1)}
    return out


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
