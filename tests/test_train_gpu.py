"""Training slice (first row): ResnetBlock2D forward + backward on the B200 kernels against torch
autograd of the oracle's resnet_block (fp32 math on the same fp16-rounded parameters and inputs).

Gates: the forward at the per-op gate; gradients -- which pass through two fp16-stored intermediate gradients (dn2 / dh1 /
dn1) -- at rel_l2 <= 5e-3 for dx and every weight gradient, 1e-2 for the small reductions (biases, norm affine
parameters, time embedding), measured ~1e-3."""
import pytest

gpu = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _block(cin, cout, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k)          # noqa: E731
    sd = {"norm1.weight": 1 + 0.2 * r(cin), "norm1.bias": 0.2 * r(cin),
          "conv1.weight": r(cout, cin, 3, 3, k=(9 * cin) ** -0.5), "conv1.bias": 0.1 * r(cout),
          "time_emb_proj.weight": r(cout, 16, k=0.25), "time_emb_proj.bias": 0.1 * r(cout),
          "norm2.weight": 1 + 0.2 * r(cout), "norm2.bias": 0.2 * r(cout),
          "conv2.weight": r(cout, cout, 3, 3, k=(9 * cout) ** -0.5), "conv2.bias": 0.1 * r(cout)}
    if cin != cout:
        sd["conv_shortcut.weight"] = r(cout, cin, 1, 1, k=cin ** -0.5)
        sd["conv_shortcut.bias"] = 0.1 * r(cout)
    # the kernels see fp16 conv weights: give the reference the same rounded values
    for k in sd:
        if k.endswith("weight") and sd[k].dim() == 4:
            sd[k] = sd[k].half().float()
    return sd, g


@gpu
@pytest.mark.parametrize("B,S,cin,cout,groups", [(2, 16, 64, 128, 8), (2, 32, 320, 320, 32), (1, 16, 640, 320, 32)])
def test_resnet_block_forward_backward_matches_torch_autograd(B, S, cin, cout, groups):
    import torch
    import torch.nn.functional as F
    from dataclasses import replace
    from oracle import uni_oracle as uo
    from uni_renderer_b200.train import ResnetBlockTrainer
    sd, g = _block(cin, cout, 7)
    x = torch.randn(B, cin, S, S, generator=g).half().float()
    temb = torch.randn(B, 16, generator=g)
    dout = (torch.randn(B, cout, S, S, generator=g) * 0.5).half().float()
    # ---- reference: torch autograd through the block exactly as oracle/uni_oracle.py::resnet_block computes it
    #      (ResnetBlock2D, models/unet_2d_blocks.py:1199), with the projected time embedding as a leaf
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    tp = (F.silu(temb) @ sd["time_emb_proj.weight"].t() + sd["time_emb_proj.bias"]).detach()
    tpr = tp.clone().requires_grad_(True)
    cfg = replace(uo.TINY, norm_num_groups=groups)
    h = F.silu(F.group_norm(xr, groups, p["norm1.weight"], p["norm1.bias"], cfg.norm_eps))
    h = F.conv2d(h, p["conv1.weight"], p["conv1.bias"], padding=1) + tpr[:, :, None, None]
    h = F.silu(F.group_norm(h, groups, p["norm2.weight"], p["norm2.bias"], cfg.norm_eps))
    h = F.conv2d(h, p["conv2.weight"], p["conv2.bias"], padding=1)
    sc = F.conv2d(xr, p["conv_shortcut.weight"], p["conv_shortcut.bias"]) if cin != cout else xr
    out_ref = sc + h
    with torch.no_grad():           # the inline restatement IS the oracle's block
        sd_o = {("blk." + k): v for k, v in sd.items()}
        assert _rel(uo.resnet_block(sd_o, "blk", x, temb, cfg), out_ref.detach()) < 1e-6
    (out_ref * dout).sum().backward()
    # ---- the kernels
    blk = ResnetBlockTrainer(sd, groups=groups, eps=cfg.norm_eps)
    xm = x.permute(0, 2, 3, 1).reshape(B * S * S, cin).half().cuda().contiguous()
    out = blk.forward(xm, tp.cuda(), B, S, S)
    grads = blk.backward(dout.permute(0, 2, 3, 1).reshape(B * S * S, cout).half().cuda().contiguous())
    torch.cuda.synchronize()
    nchw = lambda t, c: t.float().cpu().reshape(B, S, S, c).permute(0, 3, 1, 2)       # noqa: E731
    assert _rel(nchw(out, cout), out_ref.detach()) <= 1e-3
    assert _rel(nchw(grads["x"], cin), xr.grad) <= 5e-3, _rel(nchw(grads["x"], cin), xr.grad)
    assert _rel(grads["temb_proj"], tpr.grad) <= 1e-2
    for k in ("conv1.weight", "conv2.weight") + (("conv_shortcut.weight",) if cin != cout else ()):
        assert _rel(grads[k], p[k].grad) <= 5e-3, (k, _rel(grads[k], p[k].grad))
    for k in ("conv1.bias", "conv2.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias") + \
            (("conv_shortcut.bias",) if cin != cout else ()):
        assert _rel(grads[k], p[k].grad) <= 1e-2, (k, _rel(grads[k], p[k].grad))


@gpu
def test_wgrad_and_dgrad_kernels_against_torch():
    """The two conv gradient paths in isolation: dW by the tcgen05 weight-gradient kernel (3x3 with pixel splits, 1x1,
    plain linear with a ragged row count) and dX by the forward kernel on repacked weights, vs torch.autograd."""
    import torch
    import torch.nn.functional as F
    from uni_renderer_b200 import ops
    from uni_renderer_b200.ops import SEG_1x1, SEG_3x3
    from uni_renderer_b200.train import conv_wgrad, dgrad_weight
    g = torch.Generator().manual_seed(3)
    partial = torch.empty(8 << 20, device="cuda", dtype=torch.float32)
    for B, S, cin, cout, k in [(2, 32, 320, 640, 3), (1, 64, 64, 32, 3), (4, 8, 1280, 1280, 1), (2, 16, 96, 200, 1)]:
        x = torch.randn(B, cin, S, S, generator=g).half().float().requires_grad_(True)
        w = (torch.randn(cout, cin, k, k, generator=g) * (k * k * cin) ** -0.5).half().float().requires_grad_(True)
        dy = torch.randn(B, cout, S, S, generator=g).half().float()
        (F.conv2d(x, w, padding=k // 2) * dy).sum().backward()
        M = B * S * S
        xm = x.detach().permute(0, 2, 3, 1).reshape(M, cin).half().cuda().contiguous()
        dym = dy.permute(0, 2, 3, 1).reshape(M, cout).half().cuda().contiguous()
        dw, db = conv_wgrad(xm, cin, dym, cout, B=B, H=S if k == 3 else 0, W=S if k == 3 else 0, taps=k * k,
                            partial=partial, want_bias=cout % 8 == 0)
        torch.cuda.synchronize()
        assert _rel(dw, w.grad) <= 1e-3, (B, S, cin, cout, k, _rel(dw, w.grad))
        if db is not None:
            assert _rel(db, dy.sum((0, 2, 3))) <= 1e-3
        if cin % 32 == 0:
            dx = torch.empty(M, cin, device="cuda", dtype=torch.float16)
            ops.conv_gemm(None, [(dym, cout, SEG_3x3 if k == 3 else SEG_1x1)], dgrad_weight(w.detach()).cuda(), dx, M=M, N=cin,
                          B=B, H=S if k == 3 else 0, W=S if k == 3 else 0)
            torch.cuda.synchronize()
            assert _rel(dx.float().cpu().reshape(B, S, S, cin).permute(0, 3, 1, 2), x.grad) <= 1e-3


def _tblock(C, Dctx, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k)          # noqa: E731
    sd = {}
    for i in (1, 2, 3):
        sd[f"norm{i}.weight"], sd[f"norm{i}.bias"] = 1 + 0.2 * r(C), 0.2 * r(C)
    for a, kd in (("attn1", C), ("attn2", Dctx)):
        sd[f"{a}.to_q.weight"] = r(C, C, k=C ** -0.5).half().float()
        sd[f"{a}.to_k.weight"] = r(C, kd, k=kd ** -0.5).half().float()
        sd[f"{a}.to_v.weight"] = r(C, kd, k=kd ** -0.5).half().float()
        sd[f"{a}.to_out.0.weight"], sd[f"{a}.to_out.0.bias"] = r(C, C, k=C ** -0.5).half().float(), 0.1 * r(C)
    sd["ff.net.0.proj.weight"], sd["ff.net.0.proj.bias"] = r(8 * C, C, k=C ** -0.5).half().float(), 0.1 * r(8 * C)
    sd["ff.net.2.weight"], sd["ff.net.2.bias"] = r(C, 4 * C, k=(4 * C) ** -0.5).half().float(), 0.1 * r(C)
    return sd, g


@gpu
@pytest.mark.parametrize("B,N,C,heads,Lc,Dctx", [(2, 256, 128, 4, 16, 48), (1, 1024, 320, 8, 80, 768)])
def test_transformer_block_forward_backward_matches_torch_autograd(B, N, C, heads, Lc, Dctx):
    """BasicTransformerBlock (self-attention, cross-attention on the text context, GEGLU feed-forward; each behind its
    LayerNorm with a residual) forward + backward on the kernels vs torch autograd of oracle/uni_oracle.py::
    basic_transformer_block on the same fp16-rounded weights."""
    import torch
    from oracle import uni_oracle as uo
    from uni_renderer_b200.train import TransformerBlockTrainer
    sd, g = _tblock(C, Dctx, 11)
    x = torch.randn(B, N, C, generator=g).half().float()
    ctx = torch.randn(B, Lc, Dctx, generator=g).half().float()
    dout = (torch.randn(B, N, C, generator=g) * 0.5).half().float()
    p = {("blk." + k): v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    out_ref = uo.basic_transformer_block(p, "blk", xr, ctx, heads)
    (out_ref * dout).sum().backward()
    blk = TransformerBlockTrainer(sd, heads)
    out = blk.forward(x.reshape(B * N, C).half().cuda().contiguous(), ctx.reshape(B * Lc, Dctx).half().cuda().contiguous(), B)
    grads = blk.backward(dout.reshape(B * N, C).half().cuda().contiguous())
    torch.cuda.synchronize()
    assert _rel(out.reshape(B, N, C), out_ref.detach()) <= 1.5e-3
    assert _rel(grads["x"].reshape(B, N, C), xr.grad) <= 5e-3, _rel(grads["x"].reshape(B, N, C), xr.grad)
    for k, v in sd.items():
        gate = 5e-3 if (k.endswith("weight") and v.dim() == 2) else 1e-2
        assert _rel(grads[k], p["blk." + k].grad) <= gate, (k, _rel(grads[k], p["blk." + k].grad))


@gpu
@pytest.mark.parametrize("B,H,Nq,Nk,d", [(2, 4, 256, 256, 40), (1, 2, 4096, 4096, 40), (2, 8, 1024, 77, 40), (2, 2, 200, 136, 64),
                                         (2, 4, 64, 64, 8), (1, 4, 384, 77, 32), (2, 8, 1024, 1024, 80), (2, 4, 300, 77, 72)])
def test_flash_attention_backward_matches_torch_autograd(B, H, Nq, Nk, d):
    """unib200_attention_backward (csrc/attention_bwd_sm100.cu) + the forward's lse2 output against autograd of
    softmax(Q K^T / sqrt(d)) V in fp32 on the same fp16-rounded inputs (Attention / AttnProcessor2_0 of the reference's
    transformer blocks): self- and cross-attention shapes, ragged query / key counts, head dims 8 .. 80."""
    import torch
    from uni_renderer_b200 import ops
    from uni_renderer_b200.train import attention_backward
    g = torch.Generator().manual_seed(B * 1000 + Nq + Nk + d)
    C = H * d
    q = torch.randn(B * Nq, C, generator=g).half()
    k = torch.randn(B * Nk, C, generator=g).half()
    v = torch.randn(B * Nk, C, generator=g).half()
    dout = (torch.randn(B * Nq, C, generator=g) * 0.5).half()
    heads = lambda t, n: t.float().reshape(B, n, H, d).transpose(1, 2).clone().requires_grad_(True)      # noqa: E731
    qh, kh, vh = heads(q, Nq), heads(k, Nk), heads(v, Nk)
    w = torch.softmax((qh @ kh.transpose(-1, -2)) * d ** -0.5, dim=-1)
    o_ref = w @ vh
    (o_ref * dout.float().reshape(B, Nq, H, d).transpose(1, 2)).sum().backward()
    back = lambda t, n: t.transpose(1, 2).reshape(B * n, C)                                              # noqa: E731
    qc, kc, vc, dc = q.cuda(), k.cuda(), v.cuda(), dout.cuda()
    o = torch.empty(B * Nq, C, device="cuda", dtype=torch.float16)
    lse2 = torch.empty(B * H * Nq, device="cuda", dtype=torch.float32)
    ops.attention(None, qc, kc, vc, o, B=B, heads=H, Nq=Nq, Nk=Nk, d=d, lse2=lse2)
    dq, dk, dv = attention_backward(qc, kc, vc, o, dc, lse2, B=B, heads=H, Nq=Nq, Nk=Nk, d=d, scale=d ** -0.5)
    torch.cuda.synchronize()
    lse_ref = torch.logsumexp((qh @ kh.transpose(-1, -2)).detach() * d ** -0.5, -1) * 1.4426950408889634
    assert (lse2.cpu().reshape(B, H, Nq) - lse_ref).abs().max().item() <= 2e-3
    assert _rel(o, back(o_ref.detach(), Nq)) <= 1e-3
    for name, got, ref, n in (("dq", dq, qh.grad, Nq), ("dk", dk, kh.grad, Nk), ("dv", dv, vh.grad, Nk)):
        e = _rel(got, back(ref, n))
        assert e <= 3e-3, (name, e)
