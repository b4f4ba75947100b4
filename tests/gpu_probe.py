"""Per-op GPU probe: each case runs the C-ABI op on cuda:0 and compares with a torch fp32 computation on the same
fp16-rounded inputs.  `python tests/gpu_probe.py all` runs every case in its own subprocess (a device-side trap in
one case cannot poison the others) and prints one JSON line per case.  Used for bring-up and by tests/test_ops_gpu.py."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _err(got, ref):
    import torch
    got, ref = got.float(), ref.float()
    diff = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    return {"max_abs": diff.max().item(), "rel_to_max": diff.max().item() / denom,
            "rel_l2": (diff.norm() / (ref.norm() + 1e-12)).item(), "finite": bool(torch.isfinite(got).all())}


def _mk(shape, g, scale=1.0):
    import torch
    return (torch.randn(*shape, generator=g, device="cuda") * scale).half()


def case_gemm_identity():
    import torch
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    M, K, N = 256, 64, 64
    a = _mk((M, K), g)
    w = torch.eye(N, K, device="cuda").half()
    out = torch.zeros(M, N, device="cuda", dtype=torch.half)
    ops.conv_gemm(None, [(a, K, ops.SEG_1x1)], ops.pack_weight([(w, ops.SEG_1x1)]), out, M=M, N=N)
    torch.cuda.synchronize()
    r = _err(out, a)
    if r["max_abs"] > 0:
        bad = (out != a).nonzero()
        r["first_bad"] = bad[:8].tolist()
        r["sample_out"] = out[0, :8].tolist()
        r["sample_ref"] = a[0, :8].tolist()
    return r


def case_gemm_linear():
    import torch
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    res = {}
    for (M, K, N) in [(300, 320, 160), (1000, 768, 640), (128, 64, 32), (4096, 1280, 3840), (77 * 4, 768, 2560)]:
        a = _mk((M, K), g)
        w = _mk((N, K), g, K ** -0.5)
        bias = torch.randn(N, generator=g, device="cuda")
        r_ = _mk((M, N), g)
        out = torch.zeros(M, N, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(a, K, ops.SEG_1x1)], ops.pack_weight([(w, ops.SEG_1x1)]), out, M=M, N=N, bias=bias, res=r_)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + bias + r_.float()
        res[f"{M}x{K}x{N}"] = _err(out, ref)
    return res


def _conv_ref(x_nhwc, B, H, W, w, bias, stride=1):
    import torch.nn.functional as F
    x = x_nhwc.float().reshape(B, H, W, -1).permute(0, 3, 1, 2)
    y = F.conv2d(x, w.float(), bias, stride=stride, padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1).reshape(-1, w.shape[0])


def case_conv3x3():
    import torch
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    res = {}
    for (B, S, Ci, Co) in [(2, 16, 64, 64), (1, 8, 128, 32), (3, 8, 64, 64), (2, 4, 32, 64), (2, 32, 320, 640),
                           (4, 64, 320, 320), (2, 128, 64, 32)]:
        x = _mk((B * S * S, Ci), g)
        w = _mk((Co, Ci, 3, 3), g, (9 * Ci) ** -0.5)
        bias = torch.randn(Co, generator=g, device="cuda")
        out = torch.zeros(B * S * S, Co, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], ops.pack_weight([(w, ops.SEG_3x3)]), out, M=B * S * S, N=Co, B=B,
                      H=S, W=S, bias=bias)
        torch.cuda.synchronize()
        res[f"B{B}_S{S}_{Ci}->{Co}"] = _err(out, _conv_ref(x, B, S, S, w, bias))
    return res


def case_conv_variants():
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    res = {}
    # stride-2 downsample
    for (B, S, Cc) in [(2, 32, 64), (4, 16, 320)]:
        x = _mk((B * S * S, Cc), g)
        w = _mk((Cc, Cc, 3, 3), g, (9 * Cc) ** -0.5)
        bias = torch.randn(Cc, generator=g, device="cuda")
        So = S // 2
        out = torch.zeros(B * So * So, Cc, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(x, Cc, ops.SEG_3x3_S2)], ops.pack_weight([(w, ops.SEG_3x3_S2)]), out, M=B * So * So, N=Cc,
                      B=B, H=So, W=So, bias=bias)
        torch.cuda.synchronize()
        res[f"s2_B{B}_S{S}_{Cc}"] = _err(out, _conv_ref(x, B, S, S, w, bias, stride=2))
    # two-source (virtual concat) 3x3 + fused 1x1 shortcut on the raw concat + temb-style per-batch bias
    B, S, C1, C2, Co = 2, 16, 128, 64, 96
    h = _mk((B * S * S, Co), g)            # conv2 input
    x1, x2 = _mk((B * S * S, C1), g), _mk((B * S * S, C2), g)
    w2 = _mk((Co, Co, 3, 3), g, (9 * Co) ** -0.5)
    wsc = _mk((Co, C1 + C2, 1, 1), g, (C1 + C2) ** -0.5)
    biasb = torch.randn(B, Co, generator=g, device="cuda")
    out = torch.zeros(B * S * S, Co, device="cuda", dtype=torch.half)
    wp = ops.pack_weight([(w2, ops.SEG_3x3), (wsc[:, :C1], ops.SEG_1x1), (wsc[:, C1:], ops.SEG_1x1)])
    ops.conv_gemm(None, [(h, Co, ops.SEG_3x3), (x1, C1, ops.SEG_1x1), (x2, C2, ops.SEG_1x1)], wp, out, M=B * S * S,
                  N=Co, B=B, H=S, W=S, bias=biasb, bias_bstride=Co)
    torch.cuda.synchronize()
    ref = _conv_ref(h, B, S, S, w2, None) + _conv_ref(torch.cat([x1, x2], 1), B, S, S, wsc, None) \
        + biasb.repeat_interleave(S * S, 0)
    res["fused_shortcut_two_src"] = _err(out, ref)
    # column-slice sources (ld > C) and residual
    big = _mk((B * S * S, 256), g)
    xa = big[:, 64:192]
    w = _mk((64, 128, 3, 3), g, (9 * 128) ** -0.5)
    r_ = _mk((B * S * S, 64), g)
    out = torch.zeros(B * S * S, 64, device="cuda", dtype=torch.half)
    ops.conv_gemm(None, [(xa, 128, ops.SEG_3x3)], ops.pack_weight([(w, ops.SEG_3x3)]), out, M=B * S * S, N=64, B=B, H=S,
                  W=S, res=r_)
    torch.cuda.synchronize()
    res["slice_src_res"] = _err(out, _conv_ref(xa, B, S, S, w, None) + r_.float())
    # split-K on a tiny-M conv
    B, S, Ci, Co = 4, 8, 1280, 1280
    x = _mk((B * S * S, Ci), g)
    w = _mk((Co, Ci, 3, 3), g, (9 * Ci) ** -0.5)
    bias = torch.randn(Co, generator=g, device="cuda")
    r_ = _mk((B * S * S, Co), g)
    out = torch.zeros(B * S * S, Co, device="cuda", dtype=torch.half)
    partial = torch.empty(16 * B * S * S * Co, device="cuda", dtype=torch.float32)
    wp = ops.pack_weight([(w, ops.SEG_3x3)])
    ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], wp, out, M=B * S * S, N=Co, B=B, H=S, W=S, bias=bias, res=r_,
                  partial=partial)
    torch.cuda.synchronize()
    ref = _conv_ref(x, B, S, S, w, bias) + r_.float()
    res["splitk_auto"] = _err(out, ref)
    out2 = torch.zeros_like(out)
    ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], wp, out2, M=B * S * S, N=Co, B=B, H=S, W=S, bias=bias, res=r_, splits=1)
    torch.cuda.synchronize()
    res["splitk_off"] = _err(out2, ref)
    # GEGLU
    M, Cc = 512, 320
    a = _mk((M, Cc), g)
    w = _mk((8 * Cc, Cc), g, Cc ** -0.5)
    b_ = torch.randn(8 * Cc, generator=g, device="cuda")
    wi, bi = ops.pack_geglu(w, b_)
    out = torch.zeros(M, 4 * Cc, device="cuda", dtype=torch.half)
    ops.conv_gemm(None, [(a, Cc, ops.SEG_1x1)], ops.pack_weight([(wi, ops.SEG_1x1)]), out, M=M, N=8 * Cc, bias=bi,
                  flags=ops.EPI_GEGLU)
    torch.cuda.synchronize()
    y = a.float() @ w.float().t() + b_
    res["geglu"] = _err(out, y[:, :4 * Cc] * F.gelu(y[:, 4 * Cc:]))
    # NCHW fp32 output with tiny N (conv_out) and the fused scheduler update
    B, S, Ci, Co = 2, 16, 64, 28
    x = _mk((B * S * S, Ci), g)
    w = _mk((Co, Ci, 3, 3), g, (9 * Ci) ** -0.5)
    bias = torch.randn(Co, generator=g, device="cuda")
    out = torch.zeros(B, Co, S, S, device="cuda", dtype=torch.float32)
    wp = ops.pack_weight([(w, ops.SEG_3x3)])
    ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], wp, out, M=B * S * S, N=Co, B=B, H=S, W=S, bias=bias,
                  flags=ops.EPI_OUT_NCHW | ops.EPI_OUT_F32)
    torch.cuda.synchronize()
    ref = _conv_ref(x, B, S, S, w, bias).reshape(B, S, S, Co).permute(0, 3, 1, 2)
    res["nchw_f32_N28"] = _err(out, ref)
    lat = torch.randn(B, Co, S, S, generator=g, device="cuda")
    lat0 = lat.clone()
    coef = torch.tensor([[0.1, 0.2], [0.7, -0.3]], device="cuda")
    step = torch.tensor([1], device="cuda", dtype=torch.int32)
    nxt = torch.zeros(B * S * S, 32, device="cuda", dtype=torch.half)
    ops.conv_gemm(None, [(x, Ci, ops.SEG_3x3)], wp, nxt, M=B * S * S, N=Co, B=B, H=S, W=S, bias=bias,
                  flags=ops.EPI_OUT_NCHW | ops.EPI_AXPBY, axpby=coef, axpby_step=step, aux=lat, aux_out=lat,
                  axpby_first_channel=4, ldc=32)
    torch.cuda.synchronize()
    want = 0.7 * ref - 0.3 * lat0
    want[:, :4] = lat0[:, :4]
    res["axpby_latent"] = _err(lat, want)
    # the NHWC fp16 side output = raw prediction with the clean channels passed through, pad channels untouched (0)
    got_nhwc = nxt.reshape(B, S, S, 32).permute(0, 3, 1, 2)
    want_nhwc = torch.cat([lat0[:, :4], ref[:, 4:], torch.zeros(B, 4, S, S, device="cuda")], 1)
    res["axpby_nhwc_copy"] = _err(got_nhwc, want_nhwc)
    return res


def case_ln_fold():
    """LayerNorm folded into the surrounding GEMMs (include/unib200.h): producer GEMM (+bias +residual) writes row
    statistics, consumer GEMM takes the raw rows with gamma folded into its weights.  Reference: torch fp32
    layer_norm of the fp16-stored producer output followed by the linear (plain and GEGLU)."""
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    res = {}
    for (M, Cc, N2, geglu) in [(1024, 320, 960, False), (300, 64, 64, False), (256, 1280, 1280, False),
                               (512, 320, 2560, True), (96, 128, 1024, True)]:
        a = _mk((M, Cc), g)
        w1 = _mk((Cc, Cc), g, Cc ** -0.5)
        b1 = torch.randn(Cc, generator=g, device="cuda")
        r1 = _mk((M, Cc), g) + 3.0            # non-zero row means exercise the mean correction
        h = torch.zeros(M, Cc, device="cuda", dtype=torch.half)
        rs = torch.zeros(M, ops.rowstats_parts(Cc), 2, device="cuda", dtype=torch.float32)
        ops.conv_gemm(None, [(a, Cc, ops.SEG_1x1)], ops.pack_weight([(w1, ops.SEG_1x1)]), h, M=M, N=Cc, bias=b1, res=r1,
                      rowstats_out=rs)
        gamma = 1.0 + 0.3 * torch.randn(Cc, generator=g, device="cuda")
        beta = 0.2 * torch.randn(Cc, generator=g, device="cuda")
        w2 = _mk((N2, Cc), g, Cc ** -0.5)
        b2 = torch.randn(N2, generator=g, device="cuda")
        flags = 0
        if geglu:
            w2p, b2p = ops.pack_geglu(w2.float(), b2)
            flags = ops.EPI_GEGLU
        else:
            w2p, b2p = w2.float(), b2
        wf, wsum, bias2 = ops.fold_layernorm(w2p, b2p, gamma, beta)
        n_out = N2 // 2 if geglu else N2
        out = torch.zeros(M, n_out, device="cuda", dtype=torch.half)
        ops.conv_gemm(None, [(h, Cc, ops.SEG_1x1)], ops.pack_weight([(wf, ops.SEG_1x1)]), out, M=M, N=N2, bias=bias2,
                      flags=flags, ln=(rs, wsum, 1e-5, Cc))
        torch.cuda.synchronize()
        y = F.layer_norm(h.float(), (Cc,), gamma, beta, 1e-5)
        ref = y @ w2.float().t() + b2
        if geglu:
            ref = ref[:, :N2 // 2] * F.gelu(ref[:, N2 // 2:])
        res[f"M{M}_C{Cc}_N{N2}_{'geglu' if geglu else 'lin'}"] = _err(out, ref)
        # the statistics themselves: sum / sum of squares of the stored rows
        s1 = rs[:, :, 0].sum(1)
        res[f"M{M}_C{Cc}_rowsum"] = _err(s1, h.float().sum(1))
    return res


def case_norms():
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    res = {}
    scratch = torch.empty(1 << 20, device="cuda", dtype=torch.float32)
    for (B, HW, C1, C2, G, silu) in [(2, 256, 320, 0, 32, True), (4, 4096, 320, 0, 32, True), (2, 1024, 1280, 640, 32, True),
                                     (2, 256, 640, 320, 32, False), (2, 64, 32, 0, 8, True), (3, 64, 1280, 1280, 32, True)]:
        x1 = _mk((B * HW, C1), g) + 0.5
        x2 = (_mk((B * HW, C2), g) * 2 - 1) if C2 else None
        Cc = C1 + C2
        gamma = torch.randn(Cc, generator=g, device="cuda")
        beta = torch.randn(Cc, generator=g, device="cuda")
        out = torch.zeros(B * HW, Cc, device="cuda", dtype=torch.half)
        ops.groupnorm(None, x1, C1, x2, C2, gamma, beta, out, scratch, B=B, HW=HW, groups=G, eps=1e-5, silu=silu)
        torch.cuda.synchronize()
        x = torch.cat([x1, x2], 1) if C2 else x1
        xr = x.float().reshape(B, HW, Cc).permute(0, 2, 1)
        ref = F.group_norm(xr, G, gamma, beta, 1e-5)
        if silu:
            ref = F.silu(ref)
        res[f"gn_B{B}_HW{HW}_{C1}+{C2}"] = _err(out, ref.permute(0, 2, 1).reshape(B * HW, Cc))
    for (rows, Cc) in [(1000, 320), (77, 640), (4096, 1280), (50, 32), (64, 128)]:
        x = _mk((rows, Cc), g) * 3 + 1
        gamma = torch.randn(Cc, generator=g, device="cuda")
        beta = torch.randn(Cc, generator=g, device="cuda")
        y = torch.zeros_like(x)
        ops.layernorm(None, x, y, gamma, beta)
        torch.cuda.synchronize()
        res[f"ln_{rows}x{Cc}"] = _err(y, F.layer_norm(x.float(), (Cc,), gamma, beta, 1e-5))
    return res


def case_gn_fused():
    """GroupNorm statistics emitted by the GEMM epilogue (conv_gemm(gn=...)) + the single-launch apply that consumes them:
    the partial tables against torch sums of the fp32 result, and the normalised output against torch's GroupNorm of the
    STORED fp16 tensor(s) -- single source, two sources with another group size, 64- and 32-row blocks, a 3x3 conv
    producer (CTA pairs) and a 1x1 producer with residual."""
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    from uni_renderer_b200.ops import SEG_1x1, SEG_3x3
    g = torch.Generator(device="cuda").manual_seed(14)
    res = {}
    scratch = torch.empty(1 << 20, device="cuda", dtype=torch.float32)

    def producer(B, H, Cin, Cout, kind, gran, rows, with_res):
        M = B * H * H
        x = _mk((M, Cin), g)
        wt = torch.randn(Cout, Cin, *((1, 1) if kind == SEG_1x1 else (3, 3)), generator=g, device="cuda") * (
            (Cin * (1 if kind == SEG_1x1 else 9)) ** -0.5)
        bias = torch.randn(Cout, generator=g, device="cuda")
        r = _mk((M, Cout), g) if with_res else None
        out = torch.zeros(M, Cout, device="cuda", dtype=torch.half)
        part = torch.full((M // rows, Cout // gran, 2), float("nan"), device="cuda")
        ops.conv_gemm(None, [(x, Cin, kind)], ops.pack_weight([(wt, kind)]), out, M=M, N=Cout, B=B,
                      H=0 if kind == SEG_1x1 else H, W=0 if kind == SEG_1x1 else H, bias=bias, res=r,
                      gn=(part, gran, rows))
        torch.cuda.synchronize()
        xi = x.float().reshape(B, H, H, Cin).permute(0, 3, 1, 2)
        ref = F.conv2d(xi, wt.half().float(), bias, padding=0 if kind == SEG_1x1 else 1).permute(0, 2, 3, 1).reshape(M, Cout)
        if with_res:
            ref = ref + r.float()
        t = ref.reshape(M // rows, rows, Cout // gran, gran)
        ref_part = torch.stack([t.sum((1, 3)), (t * t).sum((1, 3))], -1)
        return out, part, ref, ref_part

    cases = [("c3_320_32x32", 2, 32, 320, 320, SEG_3x3, 10, 128, False),
             ("p1_640_32x32_res", 2, 32, 640, 640, SEG_1x1, 10, 128, True),
             ("c3_64_8x8_rows64", 4, 8, 32, 64, SEG_3x3, 4, 64, False),
             ("p1_128_rows32", 3, 8, 64, 128, SEG_1x1, 4, 32, True)]
    outs = {}
    for name, B, H, Cin, Cout, kind, gran, rows, wr in cases:
        out, part, ref, ref_part = producer(B, H, Cin, Cout, kind, gran, rows, wr)
        res[name + "_out"] = _err(out, ref)
        res[name + "_part"] = _err(part, ref_part)
        outs[name] = (out, part, B, H * H, Cout, gran, rows)
        # single-source GroupNorm from the partials
        G = 32 if Cout % 320 == 0 else 8
        gamma = torch.randn(Cout, generator=g, device="cuda")
        beta = torch.randn(Cout, generator=g, device="cuda")
        y = torch.zeros(B * H * H, Cout, device="cuda", dtype=torch.half)
        ops.groupnorm(None, out, Cout, None, 0, gamma, beta, y, scratch, B=B, HW=H * H, groups=G, eps=1e-5, silu=True,
                      parts=(part, None, gran, rows))
        torch.cuda.synchronize()
        xr = out.float().reshape(B, H * H, Cout).permute(0, 2, 1)
        refy = F.silu(F.group_norm(xr, G, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(B * H * H, Cout)
        res[name + "_gn"] = _err(y, refy)
    # two sources, group size 30 = 3 micro-groups of 10, the group at the seam straddles both tables
    o1, p1, B, HW, C1, gran, rows = outs["p1_640_32x32_res"]
    o2, p2 = outs["c3_320_32x32"][:2]
    Cc = C1 + 320
    gamma = torch.randn(Cc, generator=g, device="cuda")
    beta = torch.randn(Cc, generator=g, device="cuda")
    y = torch.zeros(B * HW, Cc, device="cuda", dtype=torch.half)
    ops.groupnorm(None, o1, C1, o2, 320, gamma, beta, y, scratch, B=B, HW=HW, groups=32, eps=1e-5, silu=True,
                  parts=(p1, p2, gran, rows))
    torch.cuda.synchronize()
    xr = torch.cat([o1, o2], 1).float().reshape(B, HW, Cc).permute(0, 2, 1)
    refy = F.silu(F.group_norm(xr, 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(B * HW, Cc)
    res["two_source_640+320_gn"] = _err(y, refy)
    return res


def case_upfold():
    """SEG_UP2x2: nearest-2x upsample + conv3x3 as four parity 2x2 convs on the low-resolution input, against torch's
    F.interpolate + conv2d with the ORIGINAL weights (the parity weights are fp32 sums of 1-4 taps rounded once to fp16,
    so the reference uses fp16-rounded single taps and the gate allows that one extra rounding), plus the GroupNorm
    statistics of the scattered output."""
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    from uni_renderer_b200.ops import SEG_UP2x2
    g = torch.Generator(device="cuda").manual_seed(15)
    res = {}
    for name, B, H, C, Cout, gran in [("up_1280_16to32", 2, 16, 1280, 1280, 10), ("up_640_32to64", 1, 32, 640, 640, 10),
                                      ("up_128_8to16", 3, 8, 128, 128, 4)]:
        M = B * H * H
        x = _mk((M, C), g)
        wt = torch.randn(Cout, C, 3, 3, generator=g, device="cuda") * (9 * C) ** -0.5
        bias = torch.randn(Cout, generator=g, device="cuda")
        out = torch.zeros(4 * M, Cout, device="cuda", dtype=torch.half)
        use_gn = (H * H) % 128 == 0
        part = torch.full((4 * M // 128, Cout // gran, 2), float("nan"), device="cuda") if use_gn else None
        ops.conv_gemm(None, [(x, C, SEG_UP2x2)], ops.pack_upsample_conv(wt), out, M=M, N=4 * Cout, B=B, H=H, W=H, bias=bias,
                      gn=(part, gran, 128) if use_gn else None)
        torch.cuda.synchronize()
        xi = F.interpolate(x.float().reshape(B, H, H, C).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
        ref = F.conv2d(xi, wt.half().float(), bias, padding=1).permute(0, 2, 3, 1).reshape(4 * M, Cout)
        res[name] = _err(out, ref)
        if use_gn:
            # per sample, micro-group totals must equal torch's (block order inside a sample is the kernel's business)
            nb = 4 * H * H // 128
            got = part.reshape(B, nb, Cout // gran, 2).sum(1)
            t = ref.reshape(B, 4 * H * H, Cout // gran, gran)
            res[name + "_part"] = _err(got, torch.stack([t.sum((1, 3)), (t * t).sum((1, 3))], -1))
    return res


def _attn_case(B, heads, Nq, Nk, d, g, fused_qkv):
    import torch
    from uni_renderer_b200 import ops
    Cc = heads * d
    if fused_qkv:
        qkv = _mk((B * Nq, 3 * Cc), g)
        q, k, v = qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:]
    else:
        q = _mk((B * Nq, Cc), g)
        kv = _mk((B * Nk, 2 * Cc), g)
        k, v = kv[:, :Cc], kv[:, Cc:]
    out = torch.zeros(B * Nq, Cc, device="cuda", dtype=torch.half)
    ops.attention(None, q, k, v, out, B=B, heads=heads, Nq=Nq, Nk=Nk, d=d)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Nq, heads, d).transpose(1, 2)
    kf = k.float().reshape(B, Nk, heads, d).transpose(1, 2)
    vf = v.float().reshape(B, Nk, heads, d).transpose(1, 2)
    w = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, -1)
    ref = (w @ vf).transpose(1, 2).reshape(B * Nq, Cc)
    return _err(out, ref)


def case_attention_d40():
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    return {"self_N256_d40": _attn_case(1, 2, 256, 256, 40, g, True),
            "self_N4096_d40": _attn_case(2, 8, 4096, 4096, 40, g, True),
            "cross_N1024_77_d40": _attn_case(2, 8, 1024, 77, 40, g, False),
            "self_N64_d8": _attn_case(2, 4, 64, 64, 8, g, True),
            "self_N1024_d32": _attn_case(2, 4, 1024, 1024, 32, g, True),
            # head dim 64 fills the tensor-memory budget of the P-in-TMEM softmax (2 x (S 128 + O 64 + P 64) columns);
            # ragged query / key counts exercise the padded last block
            "self_N512_d64": _attn_case(1, 2, 512, 512, 64, g, True),
            "self_N200_d48": _attn_case(1, 2, 200, 200, 48, g, True)}


def case_attention_d80():
    import torch
    g = torch.Generator(device="cuda").manual_seed(6)
    return {"self_N1024_d80": _attn_case(2, 8, 1024, 1024, 80, g, True),
            "cross_N1024_77_d80": _attn_case(2, 8, 1024, 77, 80, g, False),
            "self_N200_d64": _attn_case(1, 2, 200, 200, 64, g, True)}


def case_attention_d160():
    import torch
    g = torch.Generator(device="cuda").manual_seed(7)
    return {"self_N256_d160": _attn_case(2, 8, 256, 256, 160, g, True),
            "self_N64_d160": _attn_case(4, 8, 64, 64, 160, g, True),
            "cross_N64_77_d160": _attn_case(4, 8, 64, 77, 160, g, False)}


def case_misc():
    import math

    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(8)
    res = {}
    x = torch.randn(2, 28, 16, 16, generator=g, device="cuda")
    dst = torch.empty(2 * 256, 64, device="cuda", dtype=torch.half)
    ops.to_nhwc(None, x, dst, 64)
    torch.cuda.synchronize()
    ref = torch.zeros(2, 16, 16, 64, device="cuda")
    ref[..., :28] = x.permute(0, 2, 3, 1)
    res["to_nhwc"] = _err(dst, ref.reshape(-1, 64))
    back = torch.empty(2, 28, 16, 16, device="cuda", dtype=torch.float32)
    ops.from_nhwc(None, dst, back, B=2, Cn=28, HW=256)
    torch.cuda.synchronize()
    res["from_nhwc"] = _err(back, x.half().float())
    src = _mk((2 * 8 * 8, 64), g)
    up = torch.empty(2 * 16 * 16, 64, device="cuda", dtype=torch.half)
    ops.upsample2x(None, src, up, B=2, H=8, W=8, Cn=64)
    torch.cuda.synchronize()
    refu = F.interpolate(src.float().reshape(2, 8, 8, 64).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    res["upsample2x"] = _err(up, refu.permute(0, 2, 3, 1).reshape(-1, 64))
    t = torch.tensor([981.0, 1.0, 500.0], device="cuda")
    emb = torch.empty(3, 320, device="cuda")
    ops.timestep_sinusoid(None, t, emb, B=3, dim=320)
    torch.cuda.synchronize()
    fr = torch.exp(-math.log(10000.0) * torch.arange(160, device="cuda", dtype=torch.float32) / 160)
    ang = t[:, None] * fr[None]
    res["sinusoid"] = _err(emb, torch.cat([ang.cos(), ang.sin()], -1))
    xw = torch.randn(3, 1280, generator=g, device="cuda")
    w = _mk((2000, 1280), g, 1280 ** -0.5)
    b_ = torch.randn(2000, generator=g, device="cuda")
    y = torch.empty(3, 2000, device="cuda")
    ops.gemv(None, xw, w, b_, y, silu=True)
    torch.cuda.synchronize()
    res["gemv_silu"] = _err(y, F.silu(xw @ w.float().t() + b_))
    mo, xx = torch.randn(1000, generator=g, device="cuda"), torch.randn(1000, generator=g, device="cuda")
    coef = torch.tensor([[1.0, 2.0], [0.25, -0.5]], device="cuda")
    step = torch.tensor([1], device="cuda", dtype=torch.int32)
    o = torch.empty_like(xx)
    ops.axpby(None, mo, xx, o, coef, step)
    ops.add_int(None, step, 1)
    torch.cuda.synchronize()
    res["axpby"] = _err(o, 0.25 * mo - 0.5 * xx)
    res["add_int"] = {"max_abs": float(abs(int(step.item()) - 2)), "rel_to_max": 0.0, "rel_l2": 0.0, "finite": True}
    return res


def case_program_graph():
    import torch
    from uni_renderer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    M, K, N = 512, 320, 320
    a = _mk((M, K), g)
    w = ops.pack_weight([(_mk((N, K), g, K ** -0.5), ops.SEG_1x1)])
    mid = torch.zeros(M, N, device="cuda", dtype=torch.half)
    out = torch.zeros(M, N, device="cuda", dtype=torch.half)
    prog = ops.Program()
    ops.conv_gemm(prog, [(a, K, ops.SEG_1x1)], w, mid, M=M, N=N)
    ops.conv_gemm(prog, [(mid, N, ops.SEG_1x1)], w, out, M=M, N=N, res=a)
    prog.run()
    torch.cuda.synchronize()
    ref_mid = (a.float() @ w.float().t()).half()
    ref = ref_mid.float() @ w.float().t() + a.float()
    r = {"run": _err(out, ref), "launches": prog.num_launches}
    out.zero_()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        prog.instantiate_graph()
        prog.launch_graph()
    s.synchronize()
    torch.cuda.synchronize()
    r["graph"] = _err(out, ref)
    return r


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}


def run_all(timeout=300):
    results = {}
    for name in CASES:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=timeout)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            if p.returncode == 0 and line:
                results[name] = json.loads(line[-1])
            else:
                results[name] = {"error": (p.stderr or p.stdout)[-1500:], "returncode": p.returncode}
        except subprocess.TimeoutExpired:
            results[name] = {"error": "timeout"}
        results[name]["_seconds"] = round(time.time() - t0, 1)
        print(json.dumps({name: results[name]}), flush=True)
    return results


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        run_all()
    else:
        print(json.dumps(CASES[which]()))
