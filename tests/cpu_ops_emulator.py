"""TEST INFRASTRUCTURE: a torch-CPU emulation of the C-ABI ops' documented semantics (include/unib200.h), used to
check the HOST-SIDE wiring of recorded programs (segment order, packed-weight layout, fused shortcuts, the
GEMM-softmax-GEMM attention, the quant_conv fold) against the oracle without a GPU.

It is installed by monkeypatching `uni_renderer_b200.ops` inside a test; the product never imports it, and it is not a
fallback: nothing in uni_renderer_b200/ can reach it.  Kernel-level behaviour is covered by the `-m gpu` tests.
With it the CPU suite runs the REAL recorders (engine.StreamNet, pipeline.DualStreamSampler.plan, vae.VaeNet) of every
sampling mode and compares the result with the oracle, including a two-process gloo run of the sharded loop."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from uni_renderer_b200 import _lib as L


class FakeProgram:
    """Records closures; run() executes them in order (lanes / barriers only order concurrent work on the GPU, and a
    sequential replay is one valid schedule of the same dependency graph)."""

    def __init__(self):
        self.ops = []
        self.has_graph = False

    def keep(self, *a):
        pass

    def run(self):
        for op in self.ops:
            op()

    def lane(self, n):
        pass

    def barrier(self):
        pass

    def instantiate_graph(self):
        self.has_graph = True

    def launch_graph(self):
        self.run()

    def op_info(self):
        return []

    @property
    def num_ops(self):
        return len(self.ops)

    @property
    def num_launches(self):
        return len(self.ops)


def _submit(prog, fn):
    if prog is None:
        fn()
    else:
        prog.ops.append(fn)


def _taps(x, kind, B, H, W, C):
    """x: [B*Hin*Win, >=C] -> list of [B*H*W, C] fp32 tap matrices in the kernel's K order."""
    if kind == L.SEG_1x1:
        return [x[:, :C].float()]
    s = 1 if kind == L.SEG_3x3 else 2
    Hin, Win = H * s, W * s
    img = x[:, :C].float().reshape(B, Hin, Win, C)
    if kind == L.SEG_3x3_S2P0:
        xp = F.pad(img, (0, 0, 0, 2, 0, 2))            # bottom / right only (one pixel needed, two keeps slices easy)
    else:
        xp = F.pad(img, (0, 0, 1, 1, 1, 1))
    out = []
    for dy in range(3):
        for dx in range(3):
            out.append(xp[:, dy:dy + s * H:s, dx:dx + s * W:s].reshape(B * H * W, C))
    return out


def _gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


def conv_gemm(prog, segs, weight, out, *, M, N, B=0, H=0, W=0, bias=None, bias_bstride=0, bias_step=None,
              bias_step_stride=0, res=None, flags=0, splits=0, partial=None, axpby=None, axpby_step=None, aux=None,
              aux_out=None, axpby_first_channel=0, ldc=None, rowstats_out=None, ln=None, gn=None):
    """out = epilogue(sum_k A[m, k] * W[n, k]) with the epilogue order of csrc/gemm_sm100.cu: LayerNorm-fold
    correction, bias (per batch / per step), GEGLU gate, residual, SiLU, row statistics, store (NHWC fp16, NCHW, or the
    fused scheduler update)."""
    ktot = sum((1 if k == L.SEG_1x1 else 4 if k == L.SEG_UP2x2 else 9) * ((c + 63) // 64 * 64) for _, c, k in segs)
    assert weight.dtype == torch.float16 and tuple(weight.shape) == (N, ktot) and weight.is_contiguous()
    linear = (H == 0 or W == 0)
    bn = L.load().unib200_pick_bn(N, flags)

    def run_up():
        """SEG_UP2x2 (include/unib200.h): four parity 2x2 convs over the low-resolution input, rows scattered to the
        high-resolution output; statistics blocks are (low-res 128-row tile, parity)."""
        (t, c, _), = segs
        cout, cpad = N // 4, (c + 63) // 64 * 64
        assert flags == 0 and res is None and not bias_bstride and cout % bn == 0
        xp = F.pad(t[:, :c].float().reshape(B, H, W, c), (0, 0, 1, 1, 1, 1))
        img = torch.zeros(B, 2 * H, 2 * W, cout)
        for par in range(4):
            py, px = par >> 1, par & 1
            cols = []
            for ty in range(2):
                for tx in range(2):
                    dh, dw = ty - 1 + py, tx - 1 + px
                    cols.append(F.pad(xp[:, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W].reshape(B * H * W, c), (0, cpad - c)))
            acc = torch.cat(cols, 1) @ weight[par * cout:(par + 1) * cout].float().t()
            if bias is not None:
                acc = acc + bias.float()
            img[:, py::2, px::2] = acc.reshape(B, H, W, cout)
            if gn is not None:
                part, gran, rows = gn
                assert rows == 128 and M % 128 == 0
                tt = acc.reshape(M // 128, 128, cout // gran, gran)
                part.reshape(-1)[:(M // 128) * 4 * (cout // gran) * 2].reshape(M // 128, 4, cout // gran, 2)[:, par].copy_(
                    torch.stack([tt.sum((1, 3)), (tt * tt).sum((1, 3))], -1))
        out[:4 * M, :cout].copy_(img.reshape(4 * M, cout))

    def run():
        if segs[0][2] == L.SEG_UP2x2:
            return run_up()
        cols = []
        for t, c, kind in segs:
            assert t.dtype == torch.float16 and t.stride(1) == 1
            if linear:
                assert kind == L.SEG_1x1 and t.shape[0] >= M
                taps = [t[:M, :c].float()]
            else:
                taps = _taps(t, kind, B, H, W, c)
            cpad = (c + 63) // 64 * 64
            for tp in taps:
                cols.append(F.pad(tp, (0, cpad - c)))
        acc = torch.cat(cols, 1) @ weight.float().t()
        rpb = (M // B) if B > 0 else M
        if ln is not None:                                   # consumer of LayerNorm(x): (rowstats, wsum, eps, C)
            rs, wsum, eps, Cn = ln
            s1, s2 = rs[:M, :, 0].sum(1), rs[:M, :, 1].sum(1)
            mean = s1 / Cn
            rstd = torch.rsqrt((s2 / Cn - mean * mean).clamp_min(0) + eps)
            acc = rstd[:, None] * (acc - mean[:, None] * wsum[None, :])
        if bias is not None:
            step = int(bias_step.item()) if bias_step is not None else 0
            if bias_bstride:
                nb = (M + rpb - 1) // rpb
                bt = torch.as_strided(bias, (nb, N), (bias_bstride, 1), bias.storage_offset() + step * bias_step_stride)
                acc = acc + bt.float().repeat_interleave(rpb, 0)[:M]
            else:
                acc = acc + torch.as_strided(bias, (N,), (1,), bias.storage_offset() + step * bias_step_stride).float()
        n_out = N
        if flags & L.EPI_GEGLU:                              # per N tile: value columns | gate columns
            t = acc.reshape(M, N // bn, 2, bn // 2)
            acc = (t[:, :, 0] * _gelu_erf(t[:, :, 1])).reshape(M, N // 2)
            n_out = N // 2
        if res is not None:
            acc = acc + res[:M, :n_out].float()
        if flags & L.EPI_SILU:
            acc = F.silu(acc)
        if gn is not None:                                   # GroupNorm statistics of the output (fp32, pre-rounding)
            part, gran, rows = gn
            assert not (flags & (L.EPI_GEGLU | L.EPI_OUT_NCHW)) and M % rows == 0 and N % gran == 0 and bn % gran == 0
            t = acc.reshape(M // rows, rows, N // gran, gran)
            part.reshape(-1)[:(M // rows) * (N // gran) * 2].reshape(M // rows, N // gran, 2).copy_(
                torch.stack([t.sum((1, 3)), (t * t).sum((1, 3))], -1))
        if rowstats_out is not None:                         # consumers add the parts: put the row totals in part 0
            rv = rowstats_out.reshape(-1)[:M * 2 * ((N + bn - 1) // bn) * 2].reshape(M, -1, 2)
            rv.zero_()
            rv[:, 0, 0], rv[:, 0, 1] = acc.sum(1), (acc * acc).sum(1)
        if flags & L.EPI_OUT_NCHW:
            nb = M // rpb
            pred = acc.reshape(nb, rpb, n_out).permute(0, 2, 1)                     # [nb, N, HW]
            if flags & L.EPI_AXPBY:
                row = int(axpby_step.item()) if axpby_step is not None else 0
                c_out, c_x = axpby.reshape(-1, 2)[row].tolist()
                x = aux.reshape(nb, n_out, rpb).clone()
                keep = (torch.arange(n_out) < axpby_first_channel)[None, :, None]
                if out is not None:
                    raw = torch.where(keep, x, pred)
                    out[:M, :n_out].copy_(raw.permute(0, 2, 1).reshape(M, n_out))
                aux_out.reshape(nb, n_out, rpb).copy_(torch.where(keep, x, c_out * pred + c_x * x))
            else:
                out.reshape(nb, n_out, rpb).copy_(pred)
        else:
            out[:M, :n_out].copy_(acc)
    _submit(prog, run)


def conv_gemm_dual(prog, a, b):
    """unib200_conv_gemm_dual: two problems of identical shape in one launch = the two conv_gemm results."""
    assert a["M"] == b["M"] and a["N"] == b["N"] and len(a["segs"]) == len(b["segs"]) == 1
    for d in (a, b):
        d = dict(d)
        conv_gemm(prog, d.pop("segs"), d.pop("weight"), d.pop("out"), **d)


def t_attention_backward(q, k, v, o, dout, lse2, *, B, heads, Nq, Nk, d, scale):
    Cn = heads * d
    hs = lambda t, n: t[:, :Cn].float().reshape(B, n, heads, d).transpose(1, 2).clone().requires_grad_(True)  # noqa: E731
    qh, kh, vh = hs(q, Nq), hs(k, Nk), hs(v, Nk)
    w = torch.softmax((qh @ kh.transpose(-1, -2)) * scale, dim=-1)
    oo = w @ vh
    (oo * dout[:, :Cn].float().reshape(B, Nq, heads, d).transpose(1, 2)).sum().backward()
    back = lambda g, n: g.transpose(1, 2).reshape(B * n, Cn).half()            # noqa: E731
    return back(qh.grad, Nq), back(kh.grad, Nk), back(vh.grad, Nk)


def attention(prog, q, k, v, out, *, B, heads, Nq, Nk, d, scale=None, lse2=None):
    sc = float(scale if scale is not None else d ** -0.5)

    def run():
        for b in range(B):
            for h in range(heads):
                qs = q[b * Nq:(b + 1) * Nq, h * d:(h + 1) * d].float()
                ks = k[b * Nk:(b + 1) * Nk, h * d:(h + 1) * d].float()
                vs = v[b * Nk:(b + 1) * Nk, h * d:(h + 1) * d].float()
                sc_ = qs @ ks.t() * sc
                out[b * Nq:(b + 1) * Nq, h * d:(h + 1) * d].copy_(torch.softmax(sc_, -1) @ vs)
                if lse2 is not None:       # log2-domain log-sum-exp of every row (include/unib200.h unib200_attn_desc.lse2)
                    lse2.reshape(B, heads, Nq)[b, h].copy_(torch.logsumexp(sc_, -1) * 1.4426950408889634)
    _submit(prog, run)


def layernorm(prog, x, y, gamma, beta, eps=1e-5):
    _submit(prog, lambda: y.copy_(F.layer_norm(x.float(), (x.shape[1],), gamma, beta, eps)))


def timestep_sinusoid(prog, t, out, *, B, dim, step_idx=None, t_stride=0):
    def run():
        import math
        off = int(step_idx.item()) * t_stride if step_idx is not None else 0
        tt = t.reshape(-1)[off:off + B].float()
        half = dim // 2
        ang = tt[:, None] * torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)[None]
        out.copy_(torch.cat([torch.cos(ang), torch.sin(ang)], 1))
    _submit(prog, run)


def gemv(prog, x, w, bias, y, *, silu):
    def run():
        v = x @ w.float().t()
        if bias is not None:
            v = v + bias
        y.copy_(F.silu(v) if silu else v)
    _submit(prog, run)


def axpby(prog, model_out, x, out, coef, step_idx=None):
    def run():
        a, b = coef.reshape(-1, 2)[int(step_idx.item()) if step_idx is not None else 0].tolist()
        out.copy_(a * model_out + b * x)
    _submit(prog, run)


def unipc_step(prog, model_out, sample, last_sample, hist0, hist1, coef, step_idx=None, first_channel=0):
    def run():
        c = coef.reshape(-1, 10)[int(step_idx.item()) if step_idx is not None else 0].tolist()
        sl = (slice(None), slice(first_channel, None))
        s, h0, h1 = sample[sl].clone(), hist0[sl].clone(), hist1[sl].clone()
        x0 = c[0] * model_out[sl] + c[1] * s
        sc = c[2] * last_sample[sl] + c[3] * h0 + c[4] * h1 + c[5] * x0 if c[6] != 0 else s
        sample[sl] = c[7] * sc + c[8] * x0 + c[9] * h0
        last_sample[sl] = sc
        hist1[sl] = h0
        hist0[sl] = x0
    _submit(prog, run)


def add_int(prog, p, v):
    _submit(prog, lambda: p.add_(v))


def from_nhwc(prog, src, dst, *, B, Cn, HW):
    _submit(prog, lambda: dst.copy_(src[:, :Cn].float().reshape(B, HW, Cn).permute(0, 2, 1).reshape(dst.shape)))


def groupnorm(prog, x1, C1, x2, C2, gamma, beta, out, scratch, *, B, HW, groups, eps, silu, parts=None):
    def run():
        x = x1[:, :C1].float()
        if x2 is not None:
            x = torch.cat([x, x2[:, :C2].float()], 1)
        C = x.shape[1]
        if parts is not None:
            # statistics from the producers' epilogue partials ([B*HW / rows][C_i / gran][2]), exactly as
            # gn_apply_parts_kernel combines them: proves the host wiring hands the right tables to the right norm
            p1, p2, gran, rows = parts
            nrb = HW // rows
            mg = p1.reshape(-1)[:B * nrb * (C1 // gran) * 2].reshape(B, nrb, C1 // gran, 2).sum(1)
            if x2 is not None:
                mg = torch.cat([mg, p2.reshape(-1)[:B * nrb * (C2 // gran) * 2].reshape(B, nrb, C2 // gran, 2).sum(1)], 1)
            cpg = C // groups
            gs = mg.reshape(B, groups, cpg // gran, 2).sum(2)                 # [B, G, 2]
            mean = gs[..., 0] / (cpg * HW)
            rstd = torch.rsqrt((gs[..., 1] / (cpg * HW) - mean * mean).clamp_min(0) + eps)
            xg = x.reshape(B, HW, groups, cpg)
            y = ((xg - mean[:, None, :, None]) * rstd[:, None, :, None]).reshape(B, HW, C) * gamma + beta
            y = y.permute(0, 2, 1)
        else:
            y = F.group_norm(x.reshape(B, HW, C).permute(0, 2, 1), groups, gamma, beta, eps)
        if silu:
            y = F.silu(y)
        out.copy_(y.permute(0, 2, 1).reshape(B * HW, C))
    _submit(prog, run)


def upsample2x(prog, src, dst, *, B, H, W, Cn):
    def run():
        x = src.reshape(B, H, W, Cn)
        dst.copy_(x.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(B * 4 * H * W, Cn))
    _submit(prog, run)


def to_nhwc(prog, src, dst, Cpad):
    def run():
        B, C, H, W = src.shape
        dst.zero_()
        dst[:, :C].copy_(src.permute(0, 2, 3, 1).reshape(B * H * W, C))
    _submit(prog, run)


def softmax_rows(prog, s, *, rows, n, scale):
    def run():
        s[:rows, :n].copy_(torch.softmax(s[:rows, :n].float() * scale, dim=-1))
    _submit(prog, run)


def gaussian_sample(prog, moments, noise, out, *, scale=1.0):
    def run():
        mean, logvar = moments.chunk(2, 1)
        y = mean if noise is None else mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise
        out.copy_(y * scale)
    _submit(prog, run)


def add_f16(prog, a, b, out):
    _submit(prog, lambda: out.copy_(a.float() + b.float()))


# ---------------------------------------------------------------------------------------------------------------------
# training ops (uni_renderer_b200/train.py wrappers of the backward kernels): fp32 torch math on the fp16-stored operands,
# written independently of the kernels (autograd of the forward definition wherever that is the shortest statement)
# ---------------------------------------------------------------------------------------------------------------------
def t_conv_wgrad(x, C_in, dy, N, *, B, H, W, taps, partial=None, want_bias=True):
    M = x.shape[0]
    xf, dyf = x[:, :C_in].float(), dy[:M, :N].float()
    if taps == 1:
        dw = (dyf.t() @ xf).reshape(N, C_in, 1, 1)
    else:
        cols = _taps(x, L.SEG_3x3, B, H, W, C_in)                         # nine [M, C_in] shifted copies, (dy, dx) row-major
        dw = torch.stack([dyf.t() @ c for c in cols], -1).reshape(N, C_in, 3, 3)
    return dw.contiguous(), (dyf.sum(0) if want_bias else None)


def t_colsum(x, N):
    return x[:, :N].float().sum(0)


def t_groupnorm_backward(x, dz, gamma, beta, *, B, HW, groups, eps, silu):
    Cn = x.shape[1]
    xr = x.float().reshape(B, HW, Cn).permute(0, 2, 1).clone().requires_grad_(True)
    g, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(xr, groups, g, bt, eps)
    if silu:
        y = F.silu(y)
    (y * dz.float().reshape(B, HW, Cn).permute(0, 2, 1)).sum().backward()
    return xr.grad.permute(0, 2, 1).reshape(B * HW, Cn).half(), g.grad, bt.grad


def t_layernorm_backward(x, dy, gamma, eps=1e-5):
    xr = x.float().clone().requires_grad_(True)
    g = gamma.clone().requires_grad_(True)
    bt = torch.zeros_like(gamma).requires_grad_(True)
    (F.layer_norm(xr, (x.shape[1],), g, bt, eps) * dy.float()).sum().backward()
    return xr.grad.half(), g.grad, bt.grad


def t_geglu(proj, dout=None):
    pr = proj.float().clone().requires_grad_(True)
    a, g = pr.chunk(2, dim=-1)
    y = a * _gelu_erf(g)
    if dout is None:
        return y.detach().half()
    (y * dout.float()).sum().backward()
    return pr.grad.half()


def t_softmax_backward(p, dp, n, scale):
    pf, df = p[:, :n].float(), dp[:, :n].float()
    dp[:, :n] = (scale * pf * (df - (df * pf).sum(1, keepdim=True))).half()


def t_cvt_f32_f16(src, dst):
    dst.copy_(src.reshape(dst.shape).half())


def t_silu_f16(x, dy=None):
    v = x.float()
    sg = torch.sigmoid(v)
    return (v * sg if dy is None else dy.float() * sg * (1 + v * (1 - sg))).half()


def t_scatter2x(x, B, H, W):
    Cn = x.shape[1]
    out = torch.zeros(B, 2 * H, 2 * W, Cn, dtype=torch.float16)
    out[:, ::2, ::2] = x.reshape(B, H, W, Cn)
    return out.reshape(B * 4 * H * W, Cn)


def t_pool2x2_sum(x, B, H, W):
    Cn = x.shape[1]
    return x.float().reshape(B, H, 2, W, 2, Cn).sum((2, 4)).reshape(B * H * W, Cn).half()


def t_adamw_step(p, g, m, v, *, lr, betas, eps, weight_decay, step, grad_scale=1.0):
    gi = g * grad_scale
    m.mul_(betas[0]).add_(gi, alpha=1 - betas[0])
    v.mul_(betas[1]).addcmul_(gi, gi, value=1 - betas[1])
    p.mul_(1 - lr * weight_decay)
    p.addcdiv_(m, v.sqrt() / (1 - betas[1] ** step) ** 0.5 + eps, value=-lr / (1 - betas[0] ** step))


TRAIN_EMULATED = ("attention_backward", "conv_wgrad", "colsum", "groupnorm_backward", "layernorm_backward", "geglu", "softmax_backward",
                  "cvt_f32_f16", "silu_f16", "scatter2x", "pool2x2_sum", "adamw_step")


def install_training(monkeypatch):
    """install() + the training wrappers of uni_renderer_b200.train, and lift the trainer's CUDA gate (test only)."""
    from uni_renderer_b200 import train, trainer
    install(monkeypatch)
    for name in TRAIN_EMULATED:
        monkeypatch.setattr(train, name, globals()["t_" + name])
    monkeypatch.setattr(trainer, "_REQUIRE_CUDA", False)


EMULATED = ("conv_gemm", "conv_gemm_dual", "attention", "groupnorm", "layernorm", "upsample2x", "to_nhwc", "from_nhwc", "softmax_rows",
            "gaussian_sample", "add_f16", "add_int", "timestep_sinusoid", "gemv", "axpby", "unipc_step")


def install(monkeypatch):
    """Route uni_renderer_b200.ops through the emulator for the duration of one test."""
    from uni_renderer_b200 import ops
    for name in EMULATED:
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, "Program", FakeProgram)


def cpu_sampler(sds, cfgs, prediction_type="epsilon", scheduler="ddim", split_batch=False, temb_table=True):
    """A DualStreamSampler on the CPU behind the emulator (test only: the constructor's CUDA gate is bypassed by
    building the object by hand; call install() first)."""
    from uni_renderer_b200.engine import StreamNet, Workspace
    from uni_renderer_b200.pipeline import DualStreamSampler
    from uni_renderer_b200.scheduler import DDIMSchedule, UniPCSchedule
    s = object.__new__(DualStreamSampler)
    s.unet, s.enc, s.dec = (StreamNet(k, c, sd, "cpu") for k, c, sd in zip(("unet", "attr_enc", "attr_dec"), cfgs, sds))
    s.device = torch.device("cpu")
    s.scheduler = scheduler
    s.schedule = DDIMSchedule(prediction_type=prediction_type)
    s.unipc = UniPCSchedule(prediction_type=prediction_type)
    s.use_graph, s.split_batch, s.temb_table, s.dual_exchange = False, split_batch, temb_table, True
    s.ws, s.ws1 = Workspace("cpu"), Workspace("cpu")
    s._plans = {}
    return s
