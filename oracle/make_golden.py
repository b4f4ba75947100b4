"""Generate golden fixtures by running the REFERENCE's own model files (unmodified, from /root/reference) on CPU
under oracle/refshim (our stand-in for the un-vendored diffusers dependency).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); writes tests/golden/*.pt which
travel with the repo.  Weights are produced by oracle.uni_oracle.random_state_dict (seeded) and loaded into the
reference modules with strict=True, which also proves the oracle's key layout == the reference's state-dict layout.

    python oracle/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from models.controlnet import AttributeDecoderModel, AttributeEncoderModel, UNet2DConditionModel  # noqa: E402

from oracle import uni_oracle as uo  # noqa: E402


def build_reference(cfg_unet, cfg_enc, cfg_dec, seeds=(11, 12, 13)):
    kw = dict(block_out_channels=cfg_unet.block_out_channels, attention_head_dim=cfg_unet.num_heads,
              cross_attention_dim=cfg_unet.cross_attention_dim, norm_num_groups=cfg_unet.norm_num_groups,
              layers_per_block=cfg_unet.layers_per_block)
    unet = UNet2DConditionModel(in_channels=cfg_unet.in_channels, out_channels=cfg_unet.out_channels, **kw).eval()
    enc = AttributeEncoderModel(in_channels=cfg_enc.in_channels, **kw).eval()
    dec = AttributeDecoderModel(out_channels=cfg_dec.out_channels,
                                up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D",
                                                "CrossAttnUpBlock2D"), **kw).eval()
    sds = [uo.random_state_dict(k, c, s) for k, c, s in
           (("unet", cfg_unet, seeds[0]), ("attr_enc", cfg_enc, seeds[1]), ("attr_dec", cfg_dec, seeds[2]))]
    for m, sd in zip((unet, enc, dec), sds):
        m.load_state_dict(sd, strict=True)      # key layout + shapes must match the reference exactly
    return (unet, enc, dec), sds


def weight_digest(sd):
    return {"n": sum(v.numel() for v in sd.values()),
            "sum": float(sum(v.double().sum() for v in sd.values())),
            "abs": float(sum(v.double().abs().sum() for v in sd.values()))}


@torch.no_grad()
def run_case(name, base, B, S, t_img, t_attr, scalar_t):
    from dataclasses import replace
    cfg_unet = replace(base, in_channels=4, out_channels=4)
    cfg_enc = replace(base, in_channels=28)
    cfg_dec = replace(base, out_channels=28)
    (unet, enc, dec), sds = build_reference(cfg_unet, cfg_enc, cfg_dec)
    g = torch.Generator().manual_seed(1234)
    x_img = torch.randn(B, 4, S, S, generator=g)
    x_attr = torch.randn(B, 28, S, S, generator=g)
    ehs = torch.randn(B, 77, base.cross_attention_dim, generator=g)
    if scalar_t:
        ti, ta = t_img, t_attr
    else:
        ti, ta = torch.full((B,), t_img, dtype=torch.long), torch.full((B,), t_attr, dtype=torch.long)
    # the 3-call sequence of train/train.py:1324-1354 / models/pipeline.py:2660-2690
    d, m, raw_a, raw_a_mid = enc(x_img, ta, encoder_hidden_states=ehs, controlnet_cond=x_attr, return_dict=False)
    img_pred, raw_u, raw_u_mid, up_taps = unet(x_img, ti, encoder_hidden_states=ehs,
                                               down_block_additional_residuals=d, mid_block_additional_residual=m,
                                               return_dict=False)
    attr_pred = dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=ta, encoder_hidden_states=ehs,
                    down_block_additional_residuals=raw_u, mid_block_additional_residual=raw_u_mid,
                    return_dict=False)
    # plain UNet call without residuals (is_controlnet False path, controlnet.py:1033)
    img_plain = unet(x_img, ti, encoder_hidden_states=ehs, return_dict=False)[0]
    out = {
        "config": {"block_out_channels": base.block_out_channels, "num_heads": base.num_heads,
                   "cross_attention_dim": base.cross_attention_dim, "norm_num_groups": base.norm_num_groups,
                   "B": B, "S": S, "t_img": t_img, "t_attr": t_attr, "scalar_t": scalar_t, "seeds": (11, 12, 13)},
        "weight_digest": [weight_digest(sd) for sd in sds],
        "x_img": x_img, "x_attr": x_attr, "ehs": ehs,
        "enc_down": [t.clone() for t in d], "enc_mid": m, "enc_raw_down": [t.clone() for t in raw_a],
        "enc_raw_mid": raw_a_mid,
        "unet_sample": img_pred, "unet_raw_down": [t.clone() for t in raw_u], "unet_raw_mid": raw_u_mid,
        "unet_up_taps": [t.clone() for t in up_taps], "unet_sample_plain": img_plain,
        "dec_sample": attr_pred,
    }
    path = os.path.join(ROOT, "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(name, "->", path, os.path.getsize(path) // 1024, "KiB",
          "img_pred std %.4f attr_pred std %.4f" % (img_pred.std(), attr_pred.std()))


def sd15_checksums():
    """SD-1.5-shape single steps (BASELINE configs[0] shape: B=1, 64x64 latent, full widths) through the reference's
    own model files: t=981 (first DDIM step of a 50-step walk) and t=21 (a late step, where the scheduler update is
    dominated by the network prediction).  Stores the inputs' seed, checksums, 64 sampled values AND the full fp32
    predictions (16 K + 112 K values per case), so the GPU tests can report elementwise statistics
    (north_star's rtol/atol pass fraction) at the real shape."""
    from dataclasses import replace
    base = uo.SD15
    cfg_unet, cfg_enc, cfg_dec = replace(base), replace(base, in_channels=28), replace(base, out_channels=28)
    (unet, enc, dec), sds = build_reference(cfg_unet, cfg_enc, cfg_dec)
    B, S = 1, 64

    def one(t, seed):
        g = torch.Generator().manual_seed(seed)
        x_img = torch.randn(B, 4, S, S, generator=g)
        x_attr = torch.randn(B, 28, S, S, generator=g)
        ehs = torch.randn(B, 77, 768, generator=g)
        with torch.no_grad():
            d, m, raw_a, raw_a_mid = enc(x_img, t, encoder_hidden_states=ehs, controlnet_cond=x_attr, return_dict=False)
            img_pred, raw_u, raw_u_mid, _ = unet(x_img, t, encoder_hidden_states=ehs,
                                                 down_block_additional_residuals=d, mid_block_additional_residual=m,
                                                 return_dict=False)
            attr_pred = dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=t, encoder_hidden_states=ehs,
                            down_block_additional_residuals=raw_u, mid_block_additional_residual=raw_u_mid,
                            return_dict=False)
        idx = torch.randperm(img_pred.numel(), generator=g)[:64].clone()
        idx_a = torch.randperm(attr_pred.numel(), generator=g)[:64].clone()
        return {"config": {"B": B, "S": S, "t": t, "seed": seed, "seeds": (11, 12, 13)},
                "img_pred": {"mean": float(img_pred.mean()), "std": float(img_pred.std()), "l2": float(img_pred.norm()),
                             "idx": idx, "vals": img_pred.flatten()[idx].clone()},
                "attr_pred": {"mean": float(attr_pred.mean()), "std": float(attr_pred.std()),
                              "l2": float(attr_pred.norm()), "idx": idx_a, "vals": attr_pred.flatten()[idx_a].clone()},
                "img_pred_full": img_pred.clone(), "attr_pred_full": attr_pred.clone(),
                "raw_u_mid_l2": float(raw_u_mid.norm()), "raw_a_mid_l2": float(raw_a_mid.norm())}

    out = one(981, 1234)
    out["weight_digest"] = [weight_digest(sd) for sd in sds]
    out["n_params"] = [sum(p.numel() for p in mm.parameters()) for mm in (unet, enc, dec)]
    out["late"] = one(21, 4321)
    path = os.path.join(ROOT, "tests", "golden", "sd15_step_checksums.pt")
    torch.save(out, path)
    print("sd15 ->", path, os.path.getsize(path) // 1024, "KiB", out["n_params"], out["img_pred"]["std"],
          out["attr_pred"]["std"], out["late"]["img_pred"]["std"], out["late"]["attr_pred"]["std"])


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    if "--only-sd15" not in sys.argv:
        run_case("tiny_step_vec_t", uo.TINY, B=2, S=16, t_img=981, t_attr=981, scalar_t=False)
        run_case("tiny_step_scalar_t", uo.TINY, B=1, S=32, t_img=501, t_attr=0, scalar_t=True)
    if "--no-sd15" not in sys.argv:
        sd15_checksums()
