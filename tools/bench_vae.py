#!/usr/bin/env python
"""AutoencoderKL around the loop at BASELINE size (B=4, 512x512 images <-> 64x64 latents, SD-1.x VAE widths, random-init
weights): device time of one encode / one decode program (CUDA events), per-op-class breakdown (events around every
launch of an eager replay), algorithmic TFLOP/s against the measured tensor peak, and the image-to-image calls
(RenderPipeline.inverse_rendering / forward_rendering, 50-step DDIM) end to end with the share spent in the VAE.

    python tools/bench_vae.py [--batch 4] [--image 512] [--no-e2e] [--steps 50]

Prints one JSON line (kept under profiles/ per round)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--image", type=int, default=512)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scheduler", default="ddim", choices=["ddim", "unipc"],
                    help="unipc + --steps 20 is what the reference's shipped eval runs (eval/test_real.py:485-493)")
    ap.add_argument("--once", action="store_true", help="one decode + one encode and exit (for an ncu launch list)")
    a = ap.parse_args()
    import torch
    from bench import load_peaks
    from uni_renderer_b200 import _lib
    from uni_renderer_b200 import vae as V
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    cfg = V.VaeConfig()
    m = V.AutoencoderKL(block_out_channels=cfg.block_out_channels, down_block_types=(V._DOWN,) * 4,
                        up_block_types=(V._UP,) * 4, layers_per_block=2, norm_num_groups=32, _init_weights=False)
    m.load_state_dict(V.random_init_vae_state_dict(cfg, 21, dev))
    m = m.to(dev)
    B, S = a.batch, a.image
    h = S // 8
    g = torch.Generator(device=dev).manual_seed(4)
    z = torch.randn(B, 4, h, h, generator=g, device=dev)
    x = torch.tanh(torch.randn(B, 3, S, S, generator=g, device=dev))
    line = {"what": "AutoencoderKL (SD-1.x widths) on the B200 path", "batch": B, "image": S, "latent": h,
            "dtype": "f16 (fp32 accumulate)", "peak_tflops": peaks["tflops_sustained"], "peak_source": peaks["source"]}

    if a.once:
        m.decode(z)
        m.encode(x)
        torch.cuda.synchronize()
        return

    def timed(fn, k=5, w=3):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    for name, key, fn in (("decode", ("dec", B, h, h), lambda: m.decode(z).sample),
                          ("encode", ("enc", B, S, S), lambda: m.encode(x).latent_dist.parameters)):
        ms = timed(fn)
        P = m._progs[key]["prog"]
        info = P.op_info()
        ms_ops = P.profile(3)
        by = {}
        for (kind, fl, by_, nl), t in zip(info, ms_ops):
            d = by.setdefault(_lib.OP_NAMES[kind], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += t; d["flops"] += fl; d["bytes"] += by_; d["launches"] += nl
        flops = sum(i[1] for i in info)
        tot = sum(d["ms"] for d in by.values())
        top = sorted(((t, P.op_desc(i)) for i, t in enumerate(ms_ops)), reverse=True)[:6]
        line[name] = {
            "ms": ms, "images_per_s": B / (ms * 1e-3), "launches": P.num_launches, "flops": flops,
            "tflops": flops / (ms * 1e-3) / 1e12, "frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"],
            "eager_sum_ms": tot,
            "by_kind": {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 4), "launches": v["launches"],
                            "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                            "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None}
                        for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])},
            "top_ops": [{"ms": round(t, 4), "op": d} for t, d in top]}

    if not a.no_e2e:
        from dataclasses import replace
        from uni_renderer_b200.engine import NetConfig
        from uni_renderer_b200.models import random_init_state_dict
        from uni_renderer_b200.pipeline import DualStreamSampler
        from uni_renderer_b200.render import RenderPipeline
        nc = NetConfig(cross_attention_dim=768)
        cfgs = (replace(nc), replace(nc, in_channels=28), replace(nc, out_channels=28))
        sds = [random_init_state_dict(k, c, s, dev) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs, (11, 12, 13))]
        sampler = DualStreamSampler.from_state_dicts(*sds, *cfgs, device=dev)
        del sds
        rp = RenderPipeline(sampler, m)
        imgs = [torch.tanh(torch.randn(B, 3, S, S, generator=g, device=dev)) for _ in range(7)]
        ehs = torch.randn(B, 77, 768, generator=g, device=dev).half()
        gen = torch.Generator(device=dev).manual_seed(9)
        sch = a.scheduler
        calls = {"inverse_rendering": lambda: rp.inverse_rendering(imgs[0], imgs[1], ehs, a.steps, generator=gen,
                                                                   scheduler=sch),
                 "forward_rendering": lambda: rp.forward_rendering((0.3, 0.8), *imgs[1:7], ehs, a.steps, generator=gen,
                                                                   scheduler=sch)}
        loops = {"inverse_rendering": lambda: sampler.inverse_render(z, torch.cat([z] * 7, 1), ehs, a.steps, scheduler=sch),
                 "forward_rendering": lambda: sampler.forward_render(z, torch.cat([z] * 7, 1), ehs, a.steps, scheduler=sch)}
        for name in calls:
            ms_all = timed(calls[name], k=3, w=2)
            ms_loop = timed(loops[name], k=3, w=1)
            out = calls[name]()
            fin = all(bool(torch.isfinite(t).all()) for t in (out if isinstance(out, tuple) else (out,)))
            line[name] = {"ms": ms_all, "images_per_s": B / (ms_all * 1e-3), "loop_only_ms": ms_loop,
                          "vae_and_glue_ms": ms_all - ms_loop, "vae_share": (ms_all - ms_loop) / ms_all,
                          "denoise_steps": a.steps, "scheduler": sch, "finite": fin}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
