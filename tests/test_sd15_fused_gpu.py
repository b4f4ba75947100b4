"""The benchmarked program against the reference at its real shape: one denoising step
of the fused joint plan at SD-1.5 widths vs the predictions recorded from the reference's own model files, gated on the
recovered PREDICTION (not the latent), at the first and at a late timestep, at B = 1 and at the bench's B = 4; and the
north_star rtol 1e-3 / atol 1e-4 pass fraction reported for this path and for torch's own fp16 execution."""
import json
import os

import pytest

gpu = pytest.mark.gpu
MODEL_GATE = 3e-3           # rel_l2 on the prediction: the model-level gate of test_models_gpu.py


def _dump(name, obj):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(obj, f, indent=1)
    except OSError:
        pass


@gpu
@pytest.mark.parametrize("which,B", [("first", 1), ("late", 1), ("first", 4)])
def test_fused_joint_step_matches_reference_prediction_at_sd15_shape(which, B):
    from tests import sd15_fused_probe as P
    r = P.run_case(which, B)
    _dump(f"sd15_fused_{which}_b{B}.json", r)
    for tag in ("img", "attr"):
        assert r[tag]["rel_l2"] <= MODEL_GATE, (tag, r)
        assert r[tag]["max_abs"] <= 2e-2 * max(r[tag]["ref_absmax"], 1e-3), (tag, r)
        assert r[tag]["batch_identical"], "samples of one batch saw identical inputs and must produce identical bits"
    assert r["mask_untouched"]
    assert r["launches"] > 300


@gpu
def test_north_star_tolerance_pass_fraction_not_below_torch_fp16():
    """rtol 1e-3 / atol 1e-4 elementwise against the fp32 reference: reported for the fused plan and for torch's fp16
    execution of the same arithmetic on the same GPU; ours must not be worse (2 % slack on the fraction, 1.5x on
    rel_l2) -- the measured numbers are printed into gpurun_out/sd15_allclose.json and quoted in DESIGN.md section 5."""
    from tests import sd15_fused_probe as P
    ours = {w: P.run_case(w, 1) for w in ("first", "late")}
    ref16 = {w: P.torch_fp16_yardstick(w) for w in ("first", "late")}
    _dump("sd15_allclose.json", {"ours": ours, "torch_fp16": ref16, "rtol": P.RTOL, "atol": P.ATOL})
    for w in ("first", "late"):
        for tag in ("img", "attr"):
            assert ours[w][tag]["allclose_frac"] >= ref16[w][tag]["allclose_frac"] - 0.02, (w, tag, ours[w][tag], ref16[w][tag])
            assert ours[w][tag]["rel_l2"] <= 1.5 * ref16[w][tag]["rel_l2"] + 2e-4, (w, tag, ours[w][tag], ref16[w][tag])
