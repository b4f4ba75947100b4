"""GPU probe of the AutoencoderKL path: the ops it adds (wide-image conv, bottom/right-padded stride-2 conv, row softmax,
posterior sampling, activation-as-weight GEMMs) against torch fp32 on the same fp16-rounded inputs, and the whole
encoder / decoder against oracle/vae_oracle.py.  `python tests/vae_probe.py` prints one JSON line per case."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.gpu_probe import _err, _mk  # noqa: E402


def case_wide_conv():
    """3x3 convs on images wider than one 128-pixel tile row (VAE resolutions), incl. the fused 1x1 shortcut."""
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    res = {}
    for (B, H, W, Ci, Co) in [(1, 256, 256, 64, 32), (2, 128, 512, 32, 64), (1, 512, 512, 8, 32)]:
        x = _mk((B * H * W, Ci), g)
        w = _mk((Co, Ci, 3, 3), g, (9 * Ci) ** -0.5)
        bias = torch.randn(Co, generator=g, device="cuda")
        out = torch.zeros(B * H * W, Co, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], ops.pack_weight([(w, ops.SEG_3x3)]), out, M=B * H * W, N=Co, B=B,
                      H=H, W=W, bias=bias)
        torch.cuda.synchronize()
        xi = x.float().reshape(B, H, W, Ci).permute(0, 3, 1, 2)
        ref = F.conv2d(xi, w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
        res[f"B{B}_{H}x{W}_{Ci}->{Co}"] = _err(out, ref)
    # NCHW fp32 epilogue with N = 3 on a wide image (decoder.conv_out)
    B, H, W, Ci, Co = 2, 256, 256, 32, 3
    x = _mk((B * H * W, Ci), g)
    w = _mk((Co, Ci, 3, 3), g, (9 * Ci) ** -0.5)
    bias = torch.randn(Co, generator=g, device="cuda")
    out = torch.zeros(B, Co, H, W, device="cuda", dtype=torch.float32)
    ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], ops.pack_weight([(w, ops.SEG_3x3)]), out, M=B * H * W, N=Co, B=B, H=H,
                  W=W, bias=bias, flags=ops.EPI_OUT_NCHW | ops.EPI_OUT_F32)
    torch.cuda.synchronize()
    xi = x.float().reshape(B, H, W, Ci).permute(0, 3, 1, 2)
    res["nchw_f32_N3"] = _err(out, F.conv2d(xi, w.float(), bias, padding=1))
    return res


def case_s2p0_conv():
    """Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) + conv3x3 stride 2."""
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(12)
    res = {}
    for (B, S, Cc) in [(2, 32, 64), (1, 16, 128), (3, 8, 32), (1, 256, 32), (1, 512, 32)]:
        x = _mk((B * S * S, Cc), g)
        w = _mk((Cc, Cc, 3, 3), g, (9 * Cc) ** -0.5)
        bias = torch.randn(Cc, generator=g, device="cuda")
        So = S // 2
        out = torch.zeros(B * So * So, Cc, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(x, Cc, ops.SEG_3x3_S2P0)], ops.pack_weight([(w, ops.SEG_3x3_S2P0)]), out, M=B * So * So,
                      N=Cc, B=B, H=So, W=So, bias=bias)
        torch.cuda.synchronize()
        xi = x.float().reshape(B, S, S, Cc).permute(0, 3, 1, 2)
        ref = F.conv2d(F.pad(xi, (0, 1, 0, 1)), w.float(), bias, stride=2).permute(0, 2, 3, 1).reshape(-1, Cc)
        res[f"B{B}_S{S}_{Cc}"] = _err(out, ref)
    return res


def case_softmax_and_sample():
    import torch
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(13)
    res = {}
    for (rows, n, ld, scale) in [(256, 256, 256, 0.125), (100, 4096, 4096, 512 ** -0.5), (7, 64, 128, 1.0)]:
        s = _mk((rows, ld), g, 8.0)
        ref = torch.softmax(s[:, :n].float() * scale, dim=-1)
        keep = s[:, n:].clone()
        ops.softmax_rows(None, s, rows=rows, n=n, scale=scale)
        torch.cuda.synchronize()
        r = _err(s[:, :n], ref)
        r["rowsum_err"] = (s[:, :n].float().sum(-1) - 1).abs().max().item()
        r["pad_untouched"] = bool(torch.equal(s[:, n:], keep))
        res[f"softmax_{rows}x{n}"] = r
    mom = torch.randn(3, 8, 16, 16, generator=g, device="cuda")
    mom[:, 4:] = mom[:, 4:] * 20
    noise = torch.randn(3, 4, 16, 16, generator=g, device="cuda")
    out = torch.empty(3, 4, 16, 16, device="cuda")
    ops.gaussian_sample(None, mom, noise, out, scale=0.5)
    torch.cuda.synchronize()
    ref = (mom[:, :4] + torch.exp(0.5 * mom[:, 4:].clamp(-30, 20)) * noise) * 0.5
    res["gaussian_sample"] = _err(out, ref)
    ops.gaussian_sample(None, mom, None, out)
    torch.cuda.synchronize()
    res["gaussian_mode"] = _err(out, mom[:, :4])
    return res


def case_attention_by_gemms():
    """The mid-block attention: S = Q K^T with the keys as the [N, K] operand, row softmax, O = P V with V^T (from the
    weight-as-A GEMM) as the [N, K] operand."""
    import torch
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(14)
    res = {}
    for (T, Cc) in [(256, 64), (1024, 512)]:
        x = _mk((T, Cc), g)
        wv = _mk((Cc, Cc), g, Cc ** -0.5)
        q, k = _mk((T, Cc), g), _mk((T, Cc), g)
        bv = torch.randn(Cc, generator=g, device="cuda")
        vt = torch.zeros(Cc, T, device="cuda", dtype=torch.half)
        s = torch.zeros(T, T, device="cuda", dtype=torch.half)
        o = torch.zeros(T, Cc, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(wv, Cc, ops.SEG_1x1)], x, vt, M=Cc, N=T)
        ops.conv_gemm(None, [(q, Cc, ops.SEG_1x1)], k, s, M=T, N=T)
        ops.softmax_rows(None, s, rows=T, n=T, scale=Cc ** -0.5)
        ops.conv_gemm(None, [(s, T, ops.SEG_1x1)], vt, o, M=T, N=Cc, bias=bv)
        torch.cuda.synchronize()
        v = x.float() @ wv.float().t()
        res[f"vt_T{T}_C{Cc}"] = _err(vt, v.t())
        p = torch.softmax((q.float() @ k.float().t()) * Cc ** -0.5, dim=-1)
        res[f"attn_T{T}_C{Cc}"] = _err(o, p @ (v + bv))
    return res


def _tiny():
    from oracle import vae_oracle as vo
    from uni_renderer_b200 import vae as V
    sd = vo.random_state_dict(vo.TINY_VAE, 5)
    m = V.AutoencoderKL(block_out_channels=(32, 64, 64), down_block_types=(V._DOWN,) * 3, up_block_types=(V._UP,) * 3,
                        layers_per_block=2, norm_num_groups=8)
    m.load_state_dict(sd)
    return m.to("cuda"), sd, vo


def case_vae_decode_tiny():
    import torch
    m, sd, vo = _tiny()
    g = torch.Generator().manual_seed(1)
    res = {}
    for (B, h) in [(2, 16), (1, 64)]:             # 64x64 latent -> 256x256 image: tiles narrower than an image row
        z = torch.randn(B, 4, h, h, generator=g)
        img = m.decode(z.cuda(), return_dict=False)[0]
        img2 = m.decode(z.cuda().half()).sample
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = vo.decode(sd, vo.TINY_VAE, z)
        r = _err(img.cpu(), ref)
        r["rerun_bit_exact"] = bool(torch.equal(img, m.decode(z.cuda(), return_dict=False)[0]))
        r["fp16_in_dtype"] = str(img2.dtype)
        res[f"decode_B{B}_h{h}"] = r
    return res


def case_vae_encode_tiny():
    import torch
    m, sd, vo = _tiny()
    g = torch.Generator().manual_seed(2)
    res = {}
    for (B, S) in [(2, 64), (1, 256)]:
        x = torch.randn(B, 3, S, S, generator=g)
        dist = m.encode(x.cuda()).latent_dist
        torch.cuda.synchronize()
        with torch.no_grad():
            ref = vo.encode_moments(sd, vo.TINY_VAE, x)
        res[f"moments_B{B}_S{S}"] = _err(dist.parameters.cpu(), ref)
        res[f"mode_B{B}_S{S}"] = _err(dist.mode().cpu(), ref[:, :4])
        gen = torch.Generator(device="cuda").manual_seed(7)
        smp = dist.sample(gen)
        gen.manual_seed(7)
        noise = torch.randn(B, 4, S // 4, S // 4, generator=gen, device="cuda")
        res[f"sample_B{B}_S{S}"] = _err(smp.cpu(), vo.sample_posterior(dist.parameters.cpu(), noise.cpu()))
    return res


def case_vae_sd15_shape():
    """SD-1.x VAE widths at a reduced image size (128x128 image <-> 16x16 latent, B=1) against the oracle."""
    import torch
    from oracle import vae_oracle as vo
    from uni_renderer_b200 import vae as V
    sd = vo.random_state_dict(vo.SD15_VAE, 9)
    m = V.AutoencoderKL(block_out_channels=(128, 256, 512, 512), down_block_types=(V._DOWN,) * 4,
                        up_block_types=(V._UP,) * 4, layers_per_block=2, norm_num_groups=32)
    m.load_state_dict(sd)
    m = m.to("cuda")
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 4, 16, 16, generator=g)
    x = torch.randn(1, 3, 128, 128, generator=g)
    img = m.decode(z.cuda(), return_dict=False)[0]
    mom = m.encode(x.cuda()).latent_dist.parameters
    torch.cuda.synchronize()
    with torch.no_grad():
        return {"decode": _err(img.cpu(), vo.decode(sd, vo.SD15_VAE, z)),
                "moments": _err(mom.cpu(), vo.encode_moments(sd, vo.SD15_VAE, x))}


def case_vae_sd15_full_size():
    """BASELINE size, SD-1.x widths, one image: 64x64 latent -> 512x512 decode and 512x512 -> moments encode against
    the oracle (1M-row GEMMs, 512-pixel-wide tiles, GroupNorm over 262144 pixels, the 4096-token d = 512 attention)."""
    import torch
    from oracle import vae_oracle as vo
    from uni_renderer_b200 import vae as V
    torch.set_num_threads(os.cpu_count() or 1)
    sd = vo.random_state_dict(vo.SD15_VAE, 9)
    m = V.AutoencoderKL(block_out_channels=(128, 256, 512, 512), down_block_types=(V._DOWN,) * 4,
                        up_block_types=(V._UP,) * 4, layers_per_block=2, norm_num_groups=32)
    m.load_state_dict(sd)
    m = m.to("cuda")
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 4, 64, 64, generator=g)
    x = torch.tanh(torch.randn(1, 3, 512, 512, generator=g))
    img = m.decode(z.cuda(), return_dict=False)[0]
    mom = m.encode(x.cuda()).latent_dist.parameters
    torch.cuda.synchronize()
    with torch.no_grad():
        return {"decode_512": _err(img.cpu(), vo.decode(sd, vo.SD15_VAE, z)),
                "moments_512": _err(mom.cpu(), vo.encode_moments(sd, vo.SD15_VAE, x))}


def case_vae_full_size_timing():
    """BASELINE-sized call: B=4, 64x64 latents <-> 512x512 images, SD-1.x VAE widths, random-init weights.  Reports the
    device time of one decode and one encode (CUDA events, 2 warm + 3 timed) -- a first measurement, not a bench."""
    import torch
    from uni_renderer_b200 import vae as V
    cfg = V.VaeConfig()
    sd = V.random_init_vae_state_dict(cfg, 21)
    m = V.AutoencoderKL(block_out_channels=cfg.block_out_channels, down_block_types=(V._DOWN,) * 4,
                        up_block_types=(V._UP,) * 4, layers_per_block=2, norm_num_groups=32)
    m.load_state_dict(sd)
    m = m.to("cuda")
    g = torch.Generator(device="cuda").manual_seed(4)
    z = torch.randn(4, 4, 64, 64, generator=g, device="cuda")
    x = torch.randn(4, 3, 512, 512, generator=g, device="cuda")
    out = {}
    for name, fn in (("decode", lambda: m.decode(z).sample), ("encode", lambda: m.encode(x).latent_dist.parameters)):
        for _ in range(2):
            y = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        P = m._progs[("dec" if name == "decode" else "enc", 4, 64 if name == "decode" else 512,
                      64 if name == "decode" else 512)]["prog"]
        flops = sum(i[1] for i in P.op_info())
        ms = e0.elapsed_time(e1) / 3
        out[name] = {"ms": ms, "finite": bool(torch.isfinite(y).all()), "tflops": flops / (ms * 1e-3) / 1e12,
                     "launches": P.num_launches, "absmax": y.abs().max().item()}
    return out


def case_render_pipeline():
    """Image-to-image calls (uni_renderer_b200.render.RenderPipeline) on the tiny networks: two denoising steps, the
    same seeded noise on both sides; oracle chain = vae_oracle.encode -> uni_oracle steps + DDIM -> vae_oracle.decode.
    Also: batched / chunked decode programs agree with one-image-at-a-time calls (not bit for bit: the split-K choice
    of the small layers depends on the batch)."""
    import torch
    from tests import sampler_probe
    from oracle import uni_oracle as uo
    from uni_renderer_b200.render import RenderPipeline
    m, vsd, vo = _tiny()
    sampler, sds, cfgs = sampler_probe.tiny_setup()
    rp = RenderPipeline(sampler, m)
    B, S, steps = 2, 64, 2
    h = S // 4
    g = torch.Generator().manual_seed(31)
    imgs = [torch.tanh(torch.randn(B, 3, S, S, generator=g)) for _ in range(7)]
    ehs = torch.randn(B, 77, cfgs[0].cross_attention_dim, generator=g).half()
    sf = vo.TINY_VAE.scaling_factor
    res = {}

    def o_encode(x, noise):
        with torch.no_grad():
            return vo.sample_posterior(vo.encode_moments(vsd, vo.TINY_VAE, x), noise) * sf

    def o_loop(mode, x_img, x_attr):
        sched, sched_a = uo.DDIM(), uo.DDIM()
        ts = sched.set_timesteps(steps)
        sched_a.set_timesteps(steps)
        for i in range(steps):
            x_img, x_attr = sampler_probe.oracle_step(mode, sds, cfgs, sched, ts[i], x_img, x_attr, ehs.float(), sched_a)
        return x_img, x_attr

    # inverse rendering: image, masks -> material latents + 5 decoded attribute images
    gen = torch.Generator(device="cuda").manual_seed(77)
    out = rp.inverse_rendering(imgs[0], imgs[1], ehs, num_inference_steps=steps, generator=gen,
                               posterior_generator=gen)
    torch.cuda.synchronize()
    gen.manual_seed(77)
    n_img, n_msk = (torch.randn(B, 4, h, h, generator=gen, device="cuda").cpu() for _ in range(2))
    lat = [torch.randn(B, 4, h, h, generator=gen, device="cuda").cpu() for _ in range(6)]
    l_img, l_msk = o_encode(imgs[0], n_img), o_encode(imgs[1], n_msk)
    _, xa = o_loop("inverse", l_img, torch.cat([l_msk] + lat, 1))
    res["inv_material_latents"] = _err(out[0].cpu(), xa[:, 4:8])
    with torch.no_grad():
        for i, name in enumerate(("normal", "albedo", "spec_light", "diff_light", "env")):
            res["inv_" + name] = _err(out[1 + i].cpu(), vo.decode(vsd, vo.TINY_VAE, xa[:, 8 + 4 * i:12 + 4 * i] / sf))
    # forward rendering: 6 attribute images + material numbers -> RGB image
    gen.manual_seed(78)
    rgb = rp.forward_rendering((0.3, 0.8), *imgs[1:7], ehs, num_inference_steps=steps, generator=gen,
                               posterior_generator=gen)
    torch.cuda.synchronize()
    gen.manual_seed(78)
    noises = [torch.randn(B, 4, h, h, generator=gen, device="cuda").cpu() for _ in range(7)]
    ln, la, ls, ld_, le, lm = (o_encode(x, n) for x, n in zip(imgs[1:7], noises[:6]))
    mat = torch.empty(B, 4, h, h)
    mat[:, :2], mat[:, 2:] = 0.3 * 2 - 1, 0.8 * 2 - 1
    xi, _ = o_loop("forward", noises[6], torch.cat((lm, mat, ln, la, ls, ld_, le), 1))
    with torch.no_grad():
        res["fwd_rgb"] = _err(rgb.cpu(), vo.decode(vsd, vo.TINY_VAE, xi / sf))
    # batching is transparent: 3 images in one program == three single calls
    z = torch.randn(3, 4, h, h, generator=g).cuda()
    one = torch.cat([m.decode(z[i:i + 1]).sample for i in range(3)], 0)
    m.max_batch = 2                                   # also exercises the chunked path (2 + 1)
    res["batched_decode_vs_single"] = _err(m.decode(z).sample, one)
    return res


CASES = {"wide_conv": case_wide_conv, "s2p0_conv": case_s2p0_conv, "softmax_and_sample": case_softmax_and_sample,
         "attention_by_gemms": case_attention_by_gemms, "vae_decode_tiny": case_vae_decode_tiny,
         "vae_encode_tiny": case_vae_encode_tiny, "vae_sd15_shape": case_vae_sd15_shape, "vae_sd15_full_size": case_vae_sd15_full_size,
         "vae_full_size_timing": case_vae_full_size_timing, "render_pipeline": case_render_pipeline}


if __name__ == "__main__":
    import subprocess
    names = [a for a in sys.argv[1:] if a != "--inline"] or list(CASES)
    if "--inline" in sys.argv:    # all cases in this process (saves the per-process torch import on a fresh box)
        for n in names:
            t0 = time.time()
            try:
                r = CASES[n]()
            except Exception as e:  # noqa: BLE001
                r = {"error": f"{type(e).__name__}: {e}"}
            print(json.dumps({"case": n, "sec": round(time.time() - t0, 2), "result": r}), flush=True)
    elif len(names) == 1 and names[0] in CASES:
        t0 = time.time()
        try:
            r = CASES[names[0]]()
        except Exception as e:  # noqa: BLE001
            r = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps({"case": names[0], "sec": round(time.time() - t0, 2), "result": r}), flush=True)
    else:
        for n in names:           # one subprocess per case: a device-side trap cannot poison the others
            p = subprocess.run([sys.executable, os.path.abspath(__file__), n], capture_output=True, text=True, timeout=240)
            print(p.stdout.strip() or json.dumps({"case": n, "error": p.stderr[-2000:]}), flush=True)
