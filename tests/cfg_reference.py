"""TEST INFRASTRUCTURE: the reference's classifier-free-guidance step of each sampling loop restated on the oracle
(doubled batch, `prompt_embeds = cat([negative, positive])`, the per-loop combination rules), used by the CPU-emulator
and GPU tests of the CFG plans.  Paths relative to /root/reference."""
import torch

from oracle import uni_oracle as uo


def cfg_step(mode, sds, cfgs, sched, t, x_img, x_attr, ehs_pos, ehs_neg, g, sched_attr=None):
    """One guided denoising step.  x_img [B,4,S,S], x_attr [B,28,S,S]; returns the updated (x_img, x_attr)."""
    sa = sched_attr if sched_attr is not None else sched
    B = x_img.shape[0]
    ehs = torch.cat([ehs_neg.expand(B, -1, -1), ehs_pos], 0)               # models/pipeline.py:1445
    xi2, xa2 = torch.cat([x_img] * 2), torch.cat([x_attr] * 2)             # :1598, :2222-2235
    with torch.no_grad():
        if mode == "forward":                                              # models/pipeline.py:1586-1653
            d, m, _, _ = uo.attr_encoder_forward(sds[1], cfgs[1], 0, ehs, xa2)
            pred = uo.unet_forward(sds[0], cfgs[0], xi2, t, ehs, d, m)[0]
            cond, uncond = pred.chunk(2)                                   # :1643 (names as written in the reference)
            return sched.step(uncond + g * (cond - uncond), t, x_img), x_attr
        if mode == "inverse":                                              # models/pipeline.py:2207-2312
            _, attr = uo.dual_stream_step(*sds, *cfgs, xi2, 0, xa2, t, ehs)
            lab = attr[:, 4:]
            cond, uncond = lab.chunk(2)
            pred = cond.clone()                                            # normal .. env keep `*_pred_cond` (:2267-2285)
            pred[:, :4] = uncond[:, :4] + g * (cond[:, :4] - uncond[:, :4])   # material (:2263-2265)
            return x_img, torch.cat([x_attr[:, :4], sa.step(pred, t, x_attr[:, 4:])], 1)
        img, attr = uo.dual_stream_step(*sds, *cfgs, xi2, t, xa2, t, ehs)  # models/pipeline_new_d4p.py:1391-1453
        iu, it = img.chunk(2)                                              # `uncond, text = chunk(2)` (:1440)
        au, at = attr[:, 4:].chunk(2)
        return (sched.step(iu + g * (it - iu), t, x_img),
                torch.cat([x_attr[:, :4], sa.step(au + g * (at - au), t, x_attr[:, 4:])], 1))
