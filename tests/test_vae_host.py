"""CPU checks of the AutoencoderKL path: the drop-in surface, and the HOST-side wiring of the
recorded encoder / decoder programs executed through tests/cpu_ops_emulator.py (a torch emulation of the documented
C-ABI op semantics) against oracle/vae_oracle.py.  Kernel-level parity is in tests/test_vae_gpu.py (-m gpu)."""
import json

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as vo
from tests import cpu_ops_emulator as emu
from uni_renderer_b200 import ops
from uni_renderer_b200 import vae as V
from uni_renderer_b200.engine import Act, Workspace

TINY_KW = dict(block_out_channels=(32, 64, 64), down_block_types=(V._DOWN,) * 3, up_block_types=(V._UP,) * 3,
               layers_per_block=2, norm_num_groups=8)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("kind,stride", [(ops.SEG_3x3, 1), (ops.SEG_3x3_S2, 2), (ops.SEG_3x3_S2P0, 2), (ops.SEG_1x1, 1)])
def test_emulator_conv_semantics_match_torch(kind, stride):
    """The emulator + ops.pack_weight reproduce F.conv2d for every segment kind (3x3 pad 1, stride 2 pad 1, stride 2
    with bottom/right padding = diffusers Downsample2D(padding=0), 1x1)."""
    g = torch.Generator().manual_seed(0)
    B, S, Ci, Co = 2, 8, 24, 40
    x = torch.randn(B, Ci, S, S, generator=g).half()
    k = 1 if kind == ops.SEG_1x1 else 3
    w = (torch.randn(Co, Ci, k, k, generator=g) * 0.1).half()
    bias = torch.randn(Co, generator=g)
    So = S // stride
    xn = x.permute(0, 2, 3, 1).reshape(B * S * S, Ci).contiguous()
    out = torch.zeros(B * So * So, Co, dtype=torch.float16)
    emu.conv_gemm(None, [(xn, Ci, kind)], ops.pack_weight([(w, kind)]), out, M=B * So * So, N=Co, B=B, H=So, W=So, bias=bias)
    if kind == ops.SEG_3x3_S2P0:
        ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), bias, stride=2)
    else:
        ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=k // 2)
    assert _rel(out, ref.permute(0, 2, 3, 1).reshape(-1, Co)) < 1e-3


def _cpu_net(sd):
    """VaeNet without the CUDA gate (test only): pack on the CPU, record through the emulator."""
    net = object.__new__(V.VaeNet)
    net.cfg = V.VaeConfig(block_out_channels=(32, 64, 64), norm_num_groups=8)
    net.device = torch.device("cpu")
    net.w = {}
    net._pack(V.convert_deprecated_attention_keys(sd))
    return net


def test_decoder_wiring_matches_oracle(monkeypatch):
    emu.install(monkeypatch)
    sd = vo.random_state_dict(vo.TINY_VAE, 5)
    net = _cpu_net(sd)
    B, h = 2, 16
    g = torch.Generator().manual_seed(1)
    z = torch.randn(B, 4, h, h, generator=g)
    prog, ws = ops.Program(), Workspace("cpu")
    zin = Act(torch.zeros(B * h * h, 8, dtype=torch.float16), B, h, h, 8)
    img = torch.zeros(B, 3, 4 * h, 4 * h)
    net.rec_decoder(prog, ws, zin, img)
    ops.to_nhwc(None, z, zin.t, 8)
    prog.run()
    with torch.no_grad():
        ref = vo.decode(sd, vo.TINY_VAE, z)
    assert _rel(img, ref) < 3e-3, _rel(img, ref)


def test_encoder_wiring_matches_oracle(monkeypatch):
    emu.install(monkeypatch)
    sd = vo.random_state_dict(vo.TINY_VAE, 6)
    net = _cpu_net(sd)
    B, S = 2, 64
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 3, S, S, generator=g)
    prog, ws = ops.Program(), Workspace("cpu")
    xin = Act(torch.zeros(B * S * S, 8, dtype=torch.float16), B, S, S, 8)
    mom = torch.zeros(B, 8, S // 4, S // 4)
    net.rec_encoder(prog, ws, xin, mom)
    ops.to_nhwc(None, x, xin.t, 8)
    prog.run()
    with torch.no_grad():
        ref = vo.encode_moments(sd, vo.TINY_VAE, x)
    assert _rel(mom, ref) < 3e-3, _rel(mom, ref)
    noise = torch.randn(B, 4, S // 4, S // 4, generator=g)
    out = torch.zeros(B, 4, S // 4, S // 4)
    ops.gaussian_sample(None, mom, noise, out)
    assert _rel(out, vo.sample_posterior(ref, noise)) < 3e-3


def test_state_dict_layout_and_deprecated_attention_names():
    m = V.AutoencoderKL(**TINY_KW)
    ref = vo.random_state_dict(vo.TINY_VAE, 1)
    assert set(m.state_dict()) == set(ref)
    assert all(tuple(m.state_dict()[k].shape) == tuple(v.shape) for k, v in ref.items())
    m.load_state_dict(ref, strict=True)
    old = {}
    for k, v in ref.items():                      # LDM / diffusers < 0.20 naming, projections as linears or 1x1 convs
        for new, dep in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            if f".attentions.0.{new}." in k:
                k = k.replace(f".attentions.0.{new}.", f".attentions.0.{dep}.")
                if k.endswith("weight") and dep == "key":
                    v = v[:, :, None, None]
        old[k] = v
    m2 = V.AutoencoderKL(**TINY_KW)
    m2.load_state_dict(old, strict=True)
    assert all(torch.equal(m.state_dict()[k], m2.state_dict()[k]) for k in ref)
    assert V.vae_param_shapes(V.VaeConfig()) == vo.param_shapes(vo.SD15_VAE)     # SD-1.x VAE: 83 653 863 parameters
    assert sum(torch.Size(s).numel() for s in V.vae_param_shapes(V.VaeConfig()).values()) == 83653863


def test_surface_config_persistence_and_no_cpu_fallback(tmp_path):
    m = V.AutoencoderKL(**TINY_KW)
    assert m.config.scaling_factor == 0.18215 and len(m.config.block_out_channels) == 3      # pipeline.py:178
    assert m.dtype == torch.float16                                                         # pipeline.py:2112
    m.enable_slicing(); m.disable_slicing(); m.disable_tiling()
    with pytest.raises(NotImplementedError):
        m.enable_tiling()
    m.save_pretrained(str(tmp_path / "vae"))
    assert json.load(open(tmp_path / "vae" / "config.json"))["_class_name"] == "AutoencoderKL"
    back = V.AutoencoderKL.from_pretrained(str(tmp_path), subfolder="vae")
    assert dict(back.config) == dict(m.config)
    assert all(torch.equal(v, back.state_dict()[k]) for k, v in m.state_dict().items())
    with pytest.raises(ValueError):
        V.AutoencoderKL(down_block_types=("AttnDownEncoderBlock2D",), up_block_types=(V._UP,), block_out_channels=(64,))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.decode(torch.zeros(1, 4, 16, 16))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.encode(torch.zeros(1, 3, 64, 64))


def test_fold_quant_conv_is_exact():
    sd = vo.random_state_dict(vo.TINY_VAE, 3)
    x = torch.randn(2, 64, 8, 8)
    w, b = V.fold_quant_conv(sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], sd["quant_conv.weight"],
                             sd["quant_conv.bias"])
    two = F.conv2d(F.conv2d(x, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1),
                   sd["quant_conv.weight"], sd["quant_conv.bias"])
    torch.testing.assert_close(F.conv2d(x, w, b, padding=1), two, rtol=1e-5, atol=1e-5)


def test_module_encode_decode_api_on_emulator(monkeypatch):
    """`vae.encode(x).latent_dist.sample()/.mode()` and `vae.decode(z, return_dict=False)[0]` through the drop-in module
    (recorded programs executed by the emulator): oracle parity, the chunked path for batches above `max_batch`, the
    return forms and the dtype rules the reference's call sites rely on (models/pipeline.py:1531, 1664, 2112-2117)."""
    emu.install(monkeypatch)
    sd = vo.random_state_dict(vo.TINY_VAE, 5)
    m = V.AutoencoderKL(**TINY_KW)
    m.load_state_dict(sd)
    m.use_graph = False

    def cpu_finalize(self, device=None):            # test only: bypass the CUDA gate of the product
        if self._net is None:
            self._net = _cpu_net(dict(self.state_dict()))
            self._ws = Workspace("cpu")
        return self._net
    monkeypatch.setattr(V.AutoencoderKL, "finalize", cpu_finalize)
    g = torch.Generator().manual_seed(8)
    x = torch.tanh(torch.randn(3, 3, 32, 32, generator=g))
    z = torch.randn(3, 4, 8, 8, generator=g)
    with torch.no_grad():
        ref_m, ref_img = vo.encode_moments(sd, vo.TINY_VAE, x), vo.decode(sd, vo.TINY_VAE, z)
    out = m.encode(x)
    assert isinstance(out, V.AutoencoderKLOutput) and m.encode(x, return_dict=False)[0].parameters.shape == ref_m.shape
    dist = out.latent_dist
    assert _rel(dist.parameters, ref_m) < 3e-3 and _rel(dist.mode(), ref_m[:, :4]) < 3e-3
    assert torch.equal(dist.mean, dist.parameters[:, :4]) and torch.equal(dist.logvar, dist.parameters[:, 4:])
    g1, g2 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    smp = dist.sample(g1)
    noise = torch.randn(3, 4, 8, 8, generator=g2)
    assert _rel(smp, vo.sample_posterior(dist.parameters, noise)) < 1e-5
    img = m.decode(z, return_dict=False)[0]
    assert img.dtype == torch.float32 and _rel(img, ref_img) < 3e-3
    assert isinstance(m.decode(z), V.DecoderOutput)
    # batches above max_batch run as chunks of the recorded program and give the same result
    m.max_batch = 2
    assert _rel(m.decode(z).sample, img) < 1e-3 and _rel(m.encode(x).latent_dist.parameters, dist.parameters) < 1e-3
    assert {k[:2] for k in m._progs} == {("dec", 3), ("dec", 2), ("dec", 1), ("enc", 3), ("enc", 2), ("enc", 1)}
    with pytest.raises(ValueError):
        m.decode(torch.zeros(1, 4, 6, 6))             # latent sides must be powers of two
    full = m(x, sample_posterior=False).sample        # forward(): encode -> mode -> decode
    assert full.shape == x.shape and torch.isfinite(full).all()
