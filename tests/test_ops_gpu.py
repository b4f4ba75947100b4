"""Per-op parity on the GPU: every C-ABI kernel vs a torch fp32 computation on the same fp16-rounded inputs.
Tolerance: the only expected difference is the final fp16 rounding of the output (2^-11 relative) plus fp32
accumulation-order noise, so rel_l2 <= 5e-4 and max error <= 1e-3 of the tensor's max magnitude (fp32 outputs 1e-5)."""
import pytest

gpu = pytest.mark.gpu


def _assert_ok(res, rel_l2=5e-4, rel_max=1.5e-3):
    for name, r in res.items():
        if not isinstance(r, dict) or "rel_l2" not in r:
            continue
        assert r["finite"], name
        assert r["rel_l2"] <= rel_l2, (name, r)
        assert r["rel_to_max"] <= rel_max, (name, r)


@gpu
@pytest.mark.parametrize("case", ["gemm_linear", "conv3x3", "conv_variants", "norms", "gn_fused", "upfold", "attention_d40",
                                  "attention_d80", "attention_d160", "misc"])
def test_op_case(case):
    from tests import gpu_probe
    _assert_ok(gpu_probe.CASES[case]())


@gpu
def test_layernorm_folded_into_gemms():
    """Looser gate than the plain ops: the reference normalises in fp32 and rounds LayerNorm(x) to fp16 before the
    GEMM, the fold rounds gamma*W instead -- both are one fp16 rounding of an O(1) operand."""
    from tests import gpu_probe
    _assert_ok(gpu_probe.CASES["ln_fold"](), rel_l2=1.5e-3, rel_max=5e-3)


@gpu
def test_gemm_identity_is_exact():
    from tests import gpu_probe
    r = gpu_probe.CASES["gemm_identity"]()
    assert r["max_abs"] == 0.0, r


@gpu
def test_program_and_graph_replay():
    from tests import gpu_probe
    r = gpu_probe.CASES["program_graph"]()
    _assert_ok({"run": r["run"], "graph": r["graph"]})
    assert r["launches"] == 2
