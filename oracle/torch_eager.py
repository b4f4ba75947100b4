"""The "same-box GPU PyTorch" bar of SURVEY.md section 8d: the reference's arithmetic (our oracle restatement, which is
plain torch) executed by torch EAGER in fp16 on the B200 -- cuDNN convolutions, cuBLAS GEMMs, fused SDPA attention,
native GroupNorm / LayerNorm -- i.e. how the reference itself runs on this GPU (eval/test_real.py runs the three
modules under fp16 weights with diffusers' AttnProcessor2_0 = F.scaled_dot_product_attention).  Times the joint
dual-stream denoising step of BASELINE configs[1] (B = 4, 64x64 latents, SD-1.5 widths, random init) with CUDA events.

TEST / BASELINE INFRASTRUCTURE (it executes oracle/): never imported by the product.  bench.py runs it as a
SUBPROCESS in its `gpu_pytorch_eager` baseline leg (outside every timed region of the B200 arm), with and without
cuDNN autotuning / channels_last, and quotes the best.
    python oracle/torch_eager.py [--batch 4] [--latent 64] [--steps 5] [--warmup 3] [--cudnn-benchmark]
                                 [--channels-last] [--tiny --device cpu]
Prints one JSON line."""
import argparse
import json
import os
import sys
import time
from dataclasses import replace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--tiny", action="store_true", help="tiny widths (CPU dry run of this script)")
    ap.add_argument("--cudnn-benchmark", action="store_true")
    ap.add_argument("--channels-last", action="store_true", help="NHWC activations and conv weights (cuDNN's fast layout)")
    ap.add_argument("--train", action="store_true",
                    help="time one TRAINING step instead (train/train.py:1324-1427: fp32 master weights, fp16 autocast, "
                         "the reference's losses, backward, clip, fused AdamW) -- the torch-eager bar for uni_renderer_b200/trainer.py")
    a = ap.parse_args()
    import torch
    import torch.nn.functional as F
    from oracle import uni_oracle as uo
    from uni_renderer_b200.engine import NetConfig
    from uni_renderer_b200.models import random_init_state_dict
    dev = torch.device(a.device)
    dt = torch.float16 if dev.type == "cuda" else torch.float32
    torch.backends.cudnn.benchmark = bool(a.cudnn_benchmark)
    base_o = uo.TINY if a.tiny else uo.SD15
    nb = NetConfig(block_out_channels=base_o.block_out_channels, num_heads=base_o.num_heads,
                   cross_attention_dim=base_o.cross_attention_dim, norm_num_groups=base_o.norm_num_groups)
    cfgs_p = (replace(nb), replace(nb, in_channels=28), replace(nb, out_channels=28))
    cfgs = (replace(base_o), replace(base_o, in_channels=28), replace(base_o, out_channels=28))
    sds = [random_init_state_dict(k, c, s, dev, dtype=dt) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs_p, (11, 12, 13))]

    if a.channels_last:
        for sd in sds:
            for k, v in sd.items():
                if v.dim() == 4:
                    sd[k] = v.contiguous(memory_format=torch.channels_last)

    sin0 = uo.timestep_sinusoid
    uo.timestep_sinusoid = lambda t, dim: sin0(t.detach().float().cpu(), dim).to(device=dev, dtype=dt)

    def sdpa_attention(sd, p, x, ctx, heads):          # diffusers AttnProcessor2_0: the same math through fused SDPA
        b, n, c = x.shape
        q, k, v = uo._lin(sd, p + ".to_q", x), uo._lin(sd, p + ".to_k", ctx), uo._lin(sd, p + ".to_v", ctx)
        d = c // heads
        q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
        return uo._lin(sd, p + ".to_out.0", o)
    uo.attention = sdpa_attention

    if a.train:
        return train_main(a, torch, uo, dev, base_o, cfgs, cfgs_p, random_init_state_dict)

    B, S = a.batch, a.latent
    g = torch.Generator().manual_seed(1234)
    x_img = torch.randn(B, 4, S, S, generator=g).to(dev, dt)
    x_attr = torch.randn(B, 28, S, S, generator=g).to(dev, dt)
    if a.channels_last:
        x_img = x_img.contiguous(memory_format=torch.channels_last)
        x_attr = x_attr.contiguous(memory_format=torch.channels_last)
    ehs = torch.randn(B, 77, base_o.cross_attention_dim, generator=g).to(dev, dt)
    sched = uo.DDIM()
    ts = sched.set_timesteps(50)

    def step(i, x_img, x_attr):
        t = ts[i % 50]
        img, attr = uo.dual_stream_step(*sds, *cfgs, x_img, t, x_attr, t, ehs)
        x_attr = torch.cat([x_attr[:, :4], sched.step(attr[:, 4:], t, x_attr[:, 4:]).to(dt)], 1)
        return sched.step(img, t, x_img).to(dt), x_attr

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(a.warmup):
            x_img, x_attr = step(i, x_img, x_attr)
        sync()
        if dev.type == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for i in range(a.steps):
            x_img, x_attr = step(a.warmup + i, x_img, x_attr)
        if dev.type == "cuda":
            e1.record()
        sync()
        ms = (e0.elapsed_time(e1) if dev.type == "cuda" else (time.perf_counter() - t0) * 1e3) / a.steps
    print(json.dumps({
        "what": "torch eager fp16 (cuDNN / cuBLAS / SDPA) execution of the reference's joint dual-stream denoising step",
        "device": torch.cuda.get_device_name(0) if dev.type == "cuda" else "cpu", "torch": torch.__version__,
        "dtype": str(dt), "batch": B, "latent": S, "widths": list(base_o.block_out_channels),
        "cudnn_benchmark": bool(a.cudnn_benchmark), "channels_last": bool(a.channels_last), "steps": a.steps,
        "warmup": a.warmup,
        "ms_per_denoise_step": ms, "images_per_s_50_steps": B / (50 * ms * 1e-3),
        "finite": bool(torch.isfinite(x_img.float()).all() and torch.isfinite(x_attr.float()).all())}), flush=True)


def train_main(a, torch, uo, dev, base_o, cfgs, cfgs_p, random_init_state_dict):
    """How the reference trains on this GPU: accelerate mixed_precision="fp16" = fp32 parameters, autocast forward, scaled
    loss, torch AdamW.  SDPA attention (AttnProcessor2_0), cuDNN / cuBLAS kernels, no gradient checkpointing."""
    from uni_renderer_b200.trainer import reference_losses
    sds = [random_init_state_dict(k, c, s, dev, dtype=torch.float32)
           for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs_p, (3, 4, 5))]
    params = []
    for sd in sds:
        for k in sd:
            sd[k].requires_grad_(True)
            params.append(sd[k])
    opt = torch.optim.AdamW(params, lr=1e-5, weight_decay=1e-2, fused=dev.type == "cuda")
    B, S = a.batch, a.latent
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)                        # noqa: E731
    x_img, x_attr, ehs = r(B, 4, S, S), r(B, 28, S, S), r(B, 77, base_o.cross_attention_dim)
    t_img = torch.randint(0, 1000, (B,), generator=g).float().to(dev)
    t_attr = torch.randint(0, 1000, (B,), generator=g).float().to(dev)
    tgt_img, tgt_attr = r(B, 4, S, S), r(B, 24, S, S)
    scale = 1024.0
    losses, times = [], []
    for i in range(a.warmup + a.steps):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast(device_type=dev.type, dtype=torch.float16 if dev.type == "cuda" else torch.bfloat16):
            d, m, raw_a, raw_a_mid = uo.attr_encoder_forward(sds[1], cfgs[1], t_attr, ehs, x_attr)
            img, raw_u, raw_u_mid, _ = uo.unet_forward(sds[0], cfgs[0], x_img, t_img, ehs, d, m)
            msk = uo.attr_decoder_forward(sds[2], cfgs[2], raw_a_mid, raw_a, t_attr, ehs, raw_u, raw_u_mid)
        loss = reference_losses(img, msk, tgt_img, tgt_attr)
        (loss * scale).backward()
        for p_ in params:
            p_.grad.div_(scale)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=False)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        losses.append(float(loss.detach()))
    sec = sum(times[a.warmup:]) / a.steps
    print(json.dumps({
        "what": "torch eager training step (fp32 weights, fp16 autocast, SDPA, cuDNN / cuBLAS, fused AdamW) of the reference's "
                "3-call dual-stream step", "device": torch.cuda.get_device_name(0) if dev.type == "cuda" else "cpu",
        "torch": torch.__version__, "batch": B, "latent": S, "widths": list(base_o.block_out_channels),
        "cudnn_benchmark": bool(a.cudnn_benchmark), "seconds_per_step": sec, "images_per_s": B / sec,
        "losses": [round(x, 5) for x in losses],
        "peak_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2) if dev.type == "cuda" else None}), flush=True)


if __name__ == "__main__":
    main()
