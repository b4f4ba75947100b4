"""CPU-side checks of the C-ABI shared library: it loads, exports every symbol include/unib200.h declares, and
argument validation fails loudly (no compute is launched without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from uni_renderer_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from uni_renderer_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "unib200.h")).read()
    declared = set(re.findall(r"\b(unib200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_version_and_errors(lib):
    from uni_renderer_b200 import _lib
    assert lib.unib200_version() == 100
    d = _lib.GemmDesc()
    d.M, d.N, d.nseg = 0, 0, 0
    assert lib.unib200_conv_gemm(None, C.byref(d), None) != 0
    assert b"bad M/N/nseg" in lib.unib200_last_error()
    a = _lib.AttnDesc()
    a.d = 7
    assert lib.unib200_attention(None, C.byref(a), None) != 0
    assert b"head dim" in lib.unib200_last_error()


def test_packed_k(lib):
    from uni_renderer_b200 import _lib
    segs = (_lib.Seg * 2)()
    segs[0].C, segs[0].kind = 320, _lib.SEG_3x3
    segs[1].C, segs[1].kind = 28, _lib.SEG_1x1
    assert lib.unib200_packed_k(2, segs) == 9 * 320 + 64


def test_missing_library_raises(monkeypatch, tmp_path):
    from uni_renderer_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.Unib200Error):
        _lib.load()


@pytest.mark.gpu
def test_sampling_loop_through_the_step_level_context_with_raw_pointers():
    """after a plan exists, the `.so` alone runs the loop -- unib200_sample_loop on the plan's context with
    plain HOST pointers (numpy arrays) in and out, no torch / engine call in between; unib200_dual_step advances one
    step; unib200_load_weight / _buffer / _alloc keep context-owned device memory."""
    import ctypes as C

    import numpy as np
    import torch
    from tests import sampler_probe
    from uni_renderer_b200 import _lib
    lib = _lib.load()
    sampler, _, cfgs = sampler_probe.tiny_setup()
    B, S, steps = 2, 16, 4
    g = torch.Generator().manual_seed(5)
    x_img, x_attr = torch.randn(B, 4, S, S, generator=g), torch.randn(B, 28, S, S, generator=g)
    ehs = torch.randn(B, 77, cfgs[0].cross_attention_dim, generator=g).half()
    ref_img, ref_attr = sampler.joint_sample(x_img, x_attr, ehs.float(), num_inference_steps=steps)
    plan = sampler.plan("joint", B, S, 77, steps)
    ctx = plan.ctx.handle
    a_img, a_attr, a_ehs = x_img.numpy().copy(), x_attr.numpy().copy(), ehs.numpy().copy()
    o_img, o_attr = np.empty_like(a_img), np.empty_like(a_attr)
    stream = torch.cuda.current_stream().cuda_stream
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    assert lib.unib200_sample_loop(ctx, steps, ptr(a_img), ptr(a_attr), ptr(a_ehs), ptr(o_img), ptr(o_attr), stream) == 0
    torch.cuda.synchronize()
    assert np.array_equal(o_img, ref_img.numpy()) and np.array_equal(o_attr, ref_attr.numpy())
    # two steps by hand = a 2-step loop (the counter lives on the device)
    assert lib.unib200_sample_loop(ctx, 0, ptr(a_img), ptr(a_attr), ptr(a_ehs), None, None, stream) == 0
    assert lib.unib200_dual_step(ctx, stream) == 0 and lib.unib200_dual_step(ctx, stream) == 0
    torch.cuda.synchronize()
    two = plan.bufs["lat_img"].cpu().numpy().copy()
    assert lib.unib200_sample_loop(ctx, 2, ptr(a_img), ptr(a_attr), ptr(a_ehs), ptr(o_img), None, stream) == 0
    torch.cuda.synchronize()
    assert np.array_equal(two, o_img) and int(plan.bufs["step"].item()) == 2
    # context-owned memory
    c2 = lib.unib200_create(0, None)
    w = np.arange(64, dtype=np.float16).reshape(2, 4, 8)
    shape = (C.c_int64 * 3)(2, 4, 8)
    assert lib.unib200_load_weight(c2, b"w", ptr(w), 0, shape, 3) == 0
    nbytes = C.c_size_t()
    dptr = lib.unib200_buffer(c2, b"w", C.byref(nbytes))
    assert dptr and nbytes.value == 128 and lib.unib200_buffer(c2, b"nope", None) is None
    back = torch.zeros(64, device="cuda", dtype=torch.float16)          # read it back through an op of the library
    assert lib.unib200_add_f16(None, dptr, dptr, back.data_ptr(), 64, stream) == 0
    torch.cuda.synchronize()
    assert torch.equal(back.cpu(), 2 * torch.arange(64.).half())
    assert lib.unib200_alloc(c2, b"scratch", 1 << 20, 1) == 0 and lib.unib200_ctx_run(c2, b"step", stream) != 0
    assert b"no program" in lib.unib200_last_error()
    lib.unib200_destroy(c2)
