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
