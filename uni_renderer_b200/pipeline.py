"""Fused dual-stream sampling loops (the `for t in timesteps:` bodies of the reference's pipelines) on the B200 path.

What the reference does per denoising step (models/pipeline_new_d4p.py:1391-1453 joint; models/pipeline.py:1586-1653
forward rendering; :2627-2733 inverse rendering; train/train.py:1324-1416 for the cycle double pass):

    d, m, rawA, rawA_mid = controlnet(x_img, t_attr, ehs, controlnet_cond=x_attr28)
    img_pred, rawU, rawU_mid, _ = unet(x_img, t_img, ehs, down_block_additional_residuals=d, mid_block_additional_residual=m)
    attr_pred = controldec(rawA_mid, rawA, t_attr, ehs, down_block_additional_residuals=rawU, mid_block_additional_residual=rawU_mid)
    x_img  = scheduler_img.step(img_pred, t, x_img);   x_attr[:, 4:] = scheduler_attr.step(attr_pred[:, 4:], t, x_attr[:, 4:])

Here ONE recorded program (replayed as one CUDA graph per step) executes the whole step on static NHWC fp16 buffers:
both encoders, the two-way residual exchange as 1x1-GEMM epilogues (no torch.cat, no separate adds), both decoders and
the DDIM updates fused behind each stream's conv_out (device-side step counter indexes the timestep and coefficient
tables, so the graph is replayed without host involvement).  Work that does not change across steps is hoisted into a
per-call "setup" program: attn2 K/V of the text context for all 32 transformer blocks, the whole attribute encoder +
zero-convs for forward rendering (attr28 and t_attr = 0 are constant, pipeline.py:1455,1577-1583), the RGB encoder +
mid + decoder-side zero-convs for inverse rendering (x_img and t_img = 0 are constant, :2475; the RGB decoder output is
discarded there, :2670).  Results are identical to executing the reference's step.

No CPU / eager fallback: constructing a sampler without CUDA or without libunib200.so raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops
from .engine import Act, NetConfig, StreamNet, Workspace, pad_channels, rec_exchange_site
from .scheduler import DDIMSchedule, UniPCSchedule

SCHEDULERS = ("ddim", "unipc")

MODES = ("joint", "forward", "inverse", "cycle")
MASK_CHANNELS = 4          # the clean mask group in front of the 24 noisy attribute channels (train/train.py:1310)


@dataclass
class _Plan:
    mode: str
    B: int
    S: int
    L: int
    steps: int
    setup: ops.Program
    step: ops.Program
    bufs: Dict[str, torch.Tensor]
    flops_setup: float = 0.0
    flops_step: float = 0.0
    scheduler: str = "ddim"
    once: Optional[ops.Program] = None
    cfg: bool = False          # classifier-free guidance: the networks run on net_batch = 2 * B samples
    net_batch: int = 0
    schedule: object = None    # the DDIMSchedule / UniPCSchedule whose tables were uploaded
    timesteps: Optional[list] = None
    ctx: object = None         # ops.Context: the step-level C context that owns setup / step and runs the loop


class DualStreamSampler:
    """Owns the packed weights of the three networks and the recorded programs of the sampling loops.

    Build it from the drop-in modules (`DualStreamSampler(unet, controlnet, controldec)`), or straight from
    diffusers-layout state dicts (`from_state_dicts`).  Public entry points mirror the reference pipeline's loops:

      joint_sample(latents_img, latents_attr, prompt_embeds, num_inference_steps)   # pipeline_new_d4p.py:1287
      forward_render(latents_img, attr_latents, prompt_embeds, num_inference_steps) # pipeline.py:1368 (loop :1586)
      inverse_render(image_latents, latents_attr, prompt_embeds, num_inference_steps)  # pipeline.py:1990 (loop :2627)
      cycle_sample(...)                                                             # joint + train.py:1388-1413 pass

    Latents are [B, C, S, S] tensors on any device (host tensors are copied to the GPU; results come back on the
    device of the inputs).  `latents_attr` / `attr_latents` carry 28 channels: 4 clean mask channels + 24 attribute
    channels (pipeline.py:2647-2650).
    """

    def __init__(self, unet=None, controlnet=None, controldec=None, *, nets: Optional[Sequence[StreamNet]] = None,
                 prediction_type: str = "epsilon", device=None, use_graph: bool = True, split_batch: bool = False,
                 scheduler: str = "ddim", ddim_schedule: Optional[DDIMSchedule] = None,
                 unipc_schedule: Optional[UniPCSchedule] = None):
        if nets is None:
            if unet is None or controlnet is None or controldec is None:
                raise ValueError("need the three modules or three StreamNets")
            nets = (unet.finalize(device), controlnet.finalize(device), controldec.finalize(device))
        self.unet, self.enc, self.dec = nets
        self.device = self.unet.device
        if self.device.type != "cuda":
            raise RuntimeError("DualStreamSampler runs on CUDA (sm_100a) only")
        if scheduler not in SCHEDULERS:
            raise ValueError(f"scheduler must be one of {SCHEDULERS}")
        self.scheduler = scheduler                 # default of the sampling entry points ("ddim": BASELINE; "unipc": eval)
        # the tables default to the SD-1.x scheduler config; `set_schedules` installs the ones built from the scheduler
        # objects a caller assigned (UniRendererPipeline does that from their `.config`)
        self.schedule = ddim_schedule or DDIMSchedule(prediction_type=prediction_type)
        self.unipc = unipc_schedule or UniPCSchedule(prediction_type=prediction_type)
        self.use_graph = use_graph
        self.split_batch = split_batch
        import os
        self.temb_table = os.environ.get("UNIB200_TEMB_TABLE", "1") != "0"
        self.dual_exchange = os.environ.get("UNIB200_DUAL_EXCHANGE", "1") != "0"
        self.ws = Workspace(self.device)          # lane 0 (RGB stream)
        self.ws1 = Workspace(self.device)         # lane 1 (attribute stream): lanes run concurrently, no shared scratch
        self._plans: Dict[Tuple, _Plan] = {}

    def set_schedules(self, ddim: Optional[DDIMSchedule] = None, unipc: Optional[UniPCSchedule] = None):
        """Install other timestep / coefficient tables (plans are keyed by the tables' signature, so plans recorded for
        the previous schedule stay valid and are reused if it comes back)."""
        if ddim is not None:
            self.schedule = ddim
        if unipc is not None:
            self.unipc = unipc

    @classmethod
    def from_state_dicts(cls, sd_unet, sd_enc, sd_dec, cfg_unet: NetConfig, cfg_enc: NetConfig, cfg_dec: NetConfig,
                         device="cuda", **kw):
        nets = (StreamNet("unet", cfg_unet, sd_unet, device), StreamNet("attr_enc", cfg_enc, sd_enc, device),
                StreamNet("attr_dec", cfg_dec, sd_dec, device))
        return cls(nets=nets, **kw)

    # ------------------------------------------------------------------------------------------------------------
    # recording
    # ------------------------------------------------------------------------------------------------------------
    def plan(self, mode: str, B: int, S: int, L: int = 77, steps: int = 50, scheduler: Optional[str] = None,
             cfg: bool = False) -> _Plan:
        """cfg=True records the classifier-free-guidance variant of the loop (`guidance_scale != 0`, models/pipeline.py
        :807): the networks run on a doubled batch [first half | second half] whose text embeddings are
        cat([negative, positive]) (:1445), the two halves of each prediction are combined with per-channel-group weights
        (`_cfg_groups`) and the scheduler update runs on the combined prediction."""
        if mode not in MODES:
            raise ValueError(f"mode must be one of {MODES}")
        if cfg and mode == "cycle":
            raise NotImplementedError("the cycle double pass is a training-time construct: no guidance")
        scheduler = scheduler or self.scheduler
        if scheduler not in SCHEDULERS:
            raise ValueError(f"scheduler must be one of {SCHEDULERS}")
        if scheduler == "unipc" and mode == "cycle":
            raise NotImplementedError("the cycle double pass (a training-time construct, train/train.py:1388-1413) is "
                                      "only wired for the DDIM update")
        sched_obj = self.schedule if scheduler == "ddim" else self.unipc
        key = (mode, B, S, L, steps, scheduler, cfg, sched_obj.signature())
        if key in self._plans:
            return self._plans[key]
        Bs = B                       # samples whose latents are denoised
        B = 2 * B if cfg else B      # batch the networks see
        dev, ws, ws1 = self.device, self.ws, self.ws1
        unet, enc, dec = self.unet, self.enc, self.dec
        f32 = dict(device=dev, dtype=torch.float32)
        f16 = dict(device=dev, dtype=torch.float16)
        ci, ca = unet.cfg.in_channels, enc.cfg.in_channels
        b: Dict[str, torch.Tensor] = {}
        b["lat_img"] = torch.zeros(Bs, ci, S, S, **f32)                # x_t of the RGB stream (NCHW fp32 state)
        b["lat_attr"] = torch.zeros(Bs, ca, S, S, **f32)               # [mask | 6 attribute groups]
        b["ehs"] = torch.zeros(B * L, unet.cfg.cross_attention_dim, **f16)
        b["step"] = torch.zeros(1, device=dev, dtype=torch.int32)       # device-side step counter
        b["t_img"] = torch.zeros(steps, B, **f32)                       # per-step timestep tables
        b["t_attr"] = torch.zeros(steps, B, **f32)
        ncoef = 2 if scheduler == "ddim" else 10
        b["coef_img"] = torch.zeros(steps, ncoef, **f32)                # DDIM (c_out, c_x) / UniPC 10 scalars per step
        b["coef_attr"] = torch.zeros(steps, ncoef, **f32)
        x_img = Act(torch.zeros(B * S * S, pad_channels(ci), **f16), B, S, S, pad_channels(ci))
        x_attr = Act(torch.zeros(B * S * S, pad_channels(ca), **f16), B, S, S, pad_channels(ca))
        b["x_img"], b["x_attr"] = x_img.t, x_attr.t

        setup, step = ops.Program(), ops.Program()
        kvU = unet.rec_kv(setup, ws, b["ehs"], B, L)
        kvE = enc.rec_kv(setup, ws, b["ehs"], B, L)
        kvD = dec.rec_kv(setup, ws, b["ehs"], B, L)
        ax_img = {"coef": b["coef_img"], "step": b["step"], "latent": b["lat_img"], "first_channel": 0}
        ax_attr = {"coef": b["coef_attr"], "step": b["step"], "latent": b["lat_attr"], "first_channel": MASK_CHANNELS}

        def decode(net, wsl, mid, skips, tp, kv, ax, tag):
            """Decoder + scheduler update of one stream: DDIM fused behind conv_out, UniPC as one fused pass on the
            fp32 prediction (it needs the model-output history, so it cannot live in the conv epilogue)."""
            if cfg:
                self._rec_cfg_decode(step, b, net, wsl, mid, skips, tp, kv, L, ax, tag, mode, scheduler, Bs)
                return
            if scheduler == "ddim":
                net.rec_decoder(step, wsl, mid, skips, tp, kv, L, out_nchw=None, axpby=ax)
                return
            lat = ax["latent"]
            for nm in ("pred", "last", "h0", "h1"):
                b.setdefault(f"{nm}_{tag}", torch.zeros_like(b["lat_img" if tag == "img" else "lat_attr"]))
            # a batch-half lane passes a row slice of the latent: slice the history buffers the same way
            off = (lat.data_ptr() - b["lat_img" if tag == "img" else "lat_attr"].data_ptr()) // (lat[0].numel() * 4)
            view = (lambda t: t[off:off + lat.shape[0]])
            pred = view(b[f"pred_{tag}"])
            net.rec_decoder(step, wsl, mid, skips, tp, kv, L, out_nchw=pred)
            ops.unipc_step(step, pred, lat, view(b[f"last_{tag}"]), view(b[f"h0_{tag}"]), view(b[f"h1_{tag}"]),
                           ax["coef"], ax["step"], first_channel=ax.get("first_channel", 0))

        once = ops.Program()      # runs ONCE per plan: everything that only depends on the schedule

        def temb(net, prog, table, stepped=True):
            """Time-embedding projections: a constant-t table is computed where it is used (setup program); the
            per-step ones are tabulated for all steps up front (the timesteps of the loop are known in advance)."""
            if not stepped:
                return net.rec_temb(prog, ws, table, B)
            if not self.temb_table:       # A/B: recompute the projections inside every step
                return net.rec_temb(prog, ws, table, B, step_idx=b["step"], t_stride=B)
            return net.rec_temb_table(once, table, b["step"])

        def ingest(prog, lat, x: Act):
            """fp32 NCHW latent state -> the network's NHWC fp16 input (both halves of the doubled batch under CFG:
            `torch.cat([latents] * 2)`, models/pipeline.py:1598)."""
            rows = Bs * S * S
            for h in range(2 if cfg else 1):
                ops.to_nhwc(prog, lat, x.t[h * rows:(h + 1) * rows], x.C)

        if mode in ("joint", "cycle"):
            # The RGB stream (lane 0) and the attribute stream (lane 1) only meet at the exchange: two parallel
            # branches of the step graph, so the small-M layers of one stream fill the SMs the other leaves idle.
            step.lane(1)
            ingest(step, b["lat_attr"], x_attr)
            tpE, tpD = temb(enc, step, b["t_attr"]), temb(dec, step, b["t_attr"])
            skA, midA = enc.rec_encoder(step, ws1, x_attr, tpE, kvE, L)
            step.lane(0)
            ingest(step, b["lat_img"], x_img)
            tpU = temb(unet, step, b["t_img"])
            skU, midU = unet.rec_encoder(step, ws, x_img, tpU, kvU, L)
            step.barrier()
            if self.dual_exchange:
                # both directions of a skip site in ONE kernel (G = 2 grouped launch): 13 launches instead of 26; the
                # sites alternate between the two lanes, deepest first (the decoders consume them in that order)
                n_sk = len(skA)
                outU, outA = {}, {}
                for n_, i in enumerate(range(n_sk, -1, -1)):
                    step.lane(n_ & 1)
                    a_, u_, sfx = (midA, midU, "_mid_block") if i == n_sk else (skA[i], skU[i], f"_down_blocks.{i}")
                    outU[i], outA[i] = rec_exchange_site(step, enc, dec, sfx, a_, u_)
                step.barrier()
                dskU, dmidU = [outU[i] for i in range(n_sk)], outU[n_sk]
                dskA, dmidA = [outA[i] for i in range(n_sk)], outA[n_sk]
            else:
                dskU, dmidU = enc.rec_exchange(step, ws, skA, midA, skU, midU)      # skipU + zc_enc(skipA)
                step.lane(1)
                dskA, dmidA = dec.rec_exchange(step, ws1, skU, midU, skA, midA)     # skipA + zc_dec(skipU_raw)
            step.lane(1)
            if mode == "joint":
                decode(dec, ws1, dmidA, dskA, tpD, kvD, ax_attr, "attr")
                step.lane(0)
                decode(unet, ws, dmidU, dskU, tpU, kvU, ax_img, "img")
                step.barrier()
            else:
                step.lane(0)
                # pass 1 (train.py:1324-1355): full dual-stream step; the RGB prediction of this pass is kept as an
                # auxiliary output, the attribute prediction drives the attribute update AND feeds pass 2.
                b["img_pred_pass1"] = torch.zeros(B, unet.cfg.out_channels, S, S, **f32)
                unet.rec_decoder(step, ws, dmidU, dskU, tpU, kvU, L, out_nchw=b["img_pred_pass1"])
                step.lane(1)
                x_attr2 = Act(torch.zeros_like(x_attr.t), B, S, S, x_attr.C)
                b["x_attr_pass2"] = x_attr2.t
                dec.rec_decoder(step, ws1, dmidA, dskA, tpD, kvD, L, out_nchw=None, axpby=dict(ax_attr, nhwc=x_attr2.t))
                # pass 2 (train.py:1388-1413): attribute encoder on cat(mask, attr_pred) at t_attr = 0, then the RGB
                # stream conditioned on it.  x_img, t_img and ehs are those of pass 1, so the RGB encoder + mid of
                # pass 1 are reused (bit-identical); only the exchange and the RGB decoder are re-run.
                b["t_zero"] = torch.zeros(1, B, **f32)
                tpE0 = temb(enc, step, b["t_zero"], stepped=False)
                skA2, midA2 = enc.rec_encoder(step, ws1, x_attr2, tpE0, kvE, L)
                step.barrier()
                step.lane(0)
                dskU2, dmidU2 = enc.rec_exchange(step, ws, skA2, midA2, skU, midU)
                unet.rec_decoder(step, ws, dmidU2, dskU2, tpU, kvU, L, out_nchw=None, axpby=ax_img)
        elif mode == "forward":
            # step-invariant: attribute encoder + its 13 zero-convs (attr28, t_attr = 0, ehs constant)
            b["t_zero"] = torch.zeros(1, B, **f32)
            ingest(setup, b["lat_attr"], x_attr)
            tpE = temb(enc, setup, b["t_zero"], stepped=False)
            skA, midA = enc.rec_encoder(setup, ws, x_attr, tpE, kvE, L)
            d, m = enc.rec_exchange(setup, ws, skA, midA, [None] * len(skA), None)
            ingest(step, b["lat_img"], x_img)
            tpU = temb(unet, step, b["t_img"])
            # only one stream is live in this loop; with split_batch the two HALVES of the batch take the two lanes
            # (samples are independent).  Measured slower on B200 at B=4 (7.05 -> 7.59 ms/step: every weight is
            # streamed twice and the half-batch tiles are less efficient), so it is off by default.
            step.barrier()
            for lane, (b0, b1), wsl in self._batch_lanes(B, cfg):
                step.lane(lane)
                skU, midU = unet.rec_encoder(step, wsl, _rows(x_img, b0, b1), tpU[b0:b1], _kv_rows(kvU, b0, b1, L), L)
                dskU = [self._add(step, s_, _rows(r, b0, b1)) for s_, r in zip(skU, d)]   # controlnet.py:1078-1087
                dmidU = self._add(step, midU, _rows(m, b0, b1))                            # controlnet.py:1114-1115
                decode(unet, wsl, dmidU, dskU, tpU[b0:b1], _kv_rows(kvU, b0, b1, L),
                       dict(ax_img, latent=b["lat_img"][b0:b1]), "img")
            step.barrier()
            step.lane(0)
        else:  # inverse
            # step-invariant: RGB encoder + mid + the decoder-side zero-convs of its raw features
            b["t_zero"] = torch.zeros(1, B, **f32)
            ingest(setup, b["lat_img"], x_img)
            tpU = temb(unet, setup, b["t_zero"], stepped=False)
            skU, midU = unet.rec_encoder(setup, ws, x_img, tpU, kvU, L)
            zU, zmidU = dec.rec_exchange(setup, ws, skU, midU, [None] * len(skU), None)
            ingest(step, b["lat_attr"], x_attr)
            tpE, tpD = temb(enc, step, b["t_attr"]), temb(dec, step, b["t_attr"])
            step.barrier()
            for lane, (b0, b1), wsl in self._batch_lanes(B, cfg):  # batch halves on the two lanes, as above
                step.lane(lane)
                skA, midA = enc.rec_encoder(step, wsl, _rows(x_attr, b0, b1), tpE[b0:b1], _kv_rows(kvE, b0, b1, L), L)
                dskA = [self._add(step, s_, _rows(r, b0, b1)) for s_, r in zip(skA, zU)]   # controlnet.py:2446-2461
                dmidA = self._add(step, midA, _rows(zmidU, b0, b1))                         # controlnet.py:2476-2477
                decode(dec, wsl, dmidA, dskA, tpD[b0:b1], _kv_rows(kvD, b0, b1, L),
                       dict(ax_attr, latent=b["lat_attr"][b0:b1]), "attr")
            step.barrier()
            step.lane(0)
        ops.add_int(step, b["step"], 1)

        plan = _Plan(mode, Bs, S, L, steps, setup, step, b)
        plan.cfg, plan.net_batch = cfg, B
        plan.scheduler = scheduler
        plan.schedule = sched_obj
        plan.once = once
        plan.flops_setup = sum(i[1] for i in setup.op_info())
        plan.flops_step = sum(i[1] for i in step.op_info())
        self._upload_schedule(plan)
        once.run()                 # needs the timestep tables uploaded just above
        if self.use_graph:
            # one eager pass first (sets every kernel's function attributes outside capture), then capture on a side
            # stream: the legacy default stream cannot be captured
            setup.run()
            step.run()
            torch.cuda.synchronize(dev)
            # Programmatic dependent launch is captured into the graph as programmatic edges.  Measured on B200
            # (profiles/r2z_pdl_modes.txt): the single-lane loops (forward / inverse rendering: one dependent chain of
            # ~330 kernels per step) gain 2 % (5.68 -> 5.56 ms/step); the two-lane joint / cycle graphs do not (the
            # other lane already fills the launch gaps), so they are captured without it.
            single_lane = mode in ("forward", "inverse") and not (self.split_batch and B >= 2)
            prev_pdl = ops.set_pdl(1) if single_lane and hasattr(ops, "set_pdl") else None
            side = torch.cuda.Stream(device=dev)
            try:
                with torch.cuda.stream(side):
                    step.instantiate_graph()
                side.synchronize()
            finally:
                if prev_pdl is not None:
                    ops.set_pdl(prev_pdl)
        if self.device.type == "cuda" and hasattr(ops, "Context"):
            # hand the recorded programs to a step-level C context (include/unib200.h): from here on one call of
            # unib200_sample_loop runs setup + every denoising step, with no Python between the graph launches
            ctx = ops.Context(self.device.index or 0, use_graph=self.use_graph)
            ctx.attach("setup", setup)
            ctx.attach("step", step)
            for k in ("lat_img", "lat_attr", "ehs", "step"):
                ctx.bind(k, b[k])
            plan.ctx = ctx
        self._plans[key] = plan
        return plan

    def _batch_lanes(self, B: int, cfg: bool = False):
        """(lane, (b0, b1), workspace) per batch slice: two halves on two lanes when the batch splits evenly."""
        if self.split_batch and not cfg and B >= 2 and B % 2 == 0:
            return [(0, (0, B // 2), self.ws), (1, (B // 2, B), self.ws1)]
        return [(0, (0, B), self.ws)]

    @staticmethod
    def _cfg_groups(mode: str, tag: str):
        """(first_channel, last_channel, which) per channel group of a stream's prediction; `which` says how the two
        halves [first | second] of the doubled batch are combined with guidance scale g:
          "std"   first + g * (second - first)   joint loop: `uncond, text = chunk(2)` (pipeline_new_d4p.py:1440-1445)
          "swap"  second + g * (first - second)  forward / inverse rendering name the halves `cond, uncond = chunk(2)`
                                                 (pipeline.py:1643-1645, 2263-2265) -- mirrored as written
          "first" first                          inverse rendering keeps `*_pred_cond` for every group but material
                                                 (pipeline.py:2267-2285)"""
        if mode == "joint":
            return [(0, 4, "std")] if tag == "img" else [(MASK_CHANNELS, 28, "std")]
        if mode == "forward":
            return [(0, 4, "swap")]
        return [(MASK_CHANNELS, MASK_CHANNELS + 4, "swap"), (MASK_CHANNELS + 4, 28, "first")]

    def _rec_cfg_decode(self, step, b, net, wsl, mid, skips, tp, kv, L, ax, tag, mode, scheduler, Bs):
        """Decoder on the doubled batch -> fp32 predictions of both halves -> guided combination -> scheduler update."""
        lat = ax["latent"]
        first_channel = ax.get("first_channel", 0)
        pred2 = b.setdefault(f"pred2_{tag}", torch.zeros((2 * Bs,) + tuple(lat.shape[1:]), device=lat.device,
                                                         dtype=torch.float32))
        comb = b.setdefault(f"pred_{tag}", torch.zeros_like(lat))
        net.rec_decoder(step, wsl, mid, skips, tp, kv, L, out_nchw=pred2)
        for c0, c1, which in self._cfg_groups(mode, tag):
            w = b.setdefault(f"cfg_w_{which}", torch.zeros(1, 2, device=lat.device, dtype=torch.float32))
            for i in range(Bs):                 # a channel group of one sample is one contiguous block
                ops.axpby(step, pred2[i, c0:c1], pred2[Bs + i, c0:c1], comb[i, c0:c1], w)
        if scheduler == "ddim":
            for i in range(Bs):
                ops.axpby(step, comb[i, first_channel:], lat[i, first_channel:], lat[i, first_channel:], ax["coef"],
                          ax["step"])
        else:
            for nm in ("last", "h0", "h1"):
                b.setdefault(f"{nm}_{tag}", torch.zeros_like(lat))
            ops.unipc_step(step, comb, lat, b[f"last_{tag}"], b[f"h0_{tag}"], b[f"h1_{tag}"], ax["coef"], ax["step"],
                           first_channel=first_channel)

    def _add(self, prog, a: Act, r: Act) -> Act:
        o = Act(torch.empty_like(a.t), a.B, a.H, a.W, a.C)
        ops.add_f16(prog, a.t, r.t, o.t)
        return o

    def _upload_schedule(self, plan: _Plan):
        ts, coefs = plan.schedule.table(plan.steps)
        plan.timesteps = list(ts)
        t = torch.tensor(ts, dtype=torch.float32).reshape(-1, 1).expand(plan.steps, plan.net_batch).contiguous()
        c = torch.tensor(coefs, dtype=torch.float64).to(torch.float32)
        b = plan.bufs
        # joint/cycle: both streams walk the same timesteps; forward: t_attr = 0; inverse: t_img = 0 (hoisted)
        b["t_img"].copy_(t)
        b["t_attr"].copy_(t)
        b["coef_img"].copy_(c)
        b["coef_attr"].copy_(c)

    # ------------------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------------------
    def load_inputs(self, plan: _Plan, latents_img, latents_attr, prompt_embeds, negative_prompt_embeds=None,
                    guidance_scale: float = 0.0):
        """Copy one batch of inputs (any device / float dtype; pinned host memory makes this an async H2D) into
        the plan's static buffers.  A CFG plan also takes the negative embeddings (first half of the doubled batch,
        `torch.cat([negative_prompt_embeds, prompt_embeds])`, models/pipeline.py:1445) and the guidance scale."""
        b = plan.bufs
        if plan.cfg:
            if negative_prompt_embeds is None:
                raise ValueError("a classifier-free-guidance plan needs negative_prompt_embeds")
            rows = plan.B * plan.L
            neg = negative_prompt_embeds
            if neg.shape[0] == 1 and plan.B > 1:        # the reference only ever builds one negative row (:372-430)
                neg = neg.expand(plan.B, -1, -1)
            b["ehs"][:rows].copy_(neg.reshape(rows, -1), non_blocking=True)
            b["ehs"][rows:].copy_(prompt_embeds.reshape(rows, -1), non_blocking=True)
            g = float(guidance_scale)
            for which, w in (("std", (1.0 - g, g)), ("swap", (g, 1.0 - g)), ("first", (1.0, 0.0))):
                if f"cfg_w_{which}" in b:
                    b[f"cfg_w_{which}"].copy_(torch.tensor([w], dtype=torch.float32))
        elif negative_prompt_embeds is not None:
            raise ValueError("negative_prompt_embeds given to a plan recorded without guidance")
        if tuple(latents_img.shape) != tuple(b["lat_img"].shape) or tuple(latents_attr.shape) != tuple(b["lat_attr"].shape):
            raise ValueError(f"latent shapes {tuple(latents_img.shape)} / {tuple(latents_attr.shape)} do not match "
                             f"the plan {tuple(b['lat_img'].shape)} / {tuple(b['lat_attr'].shape)}")
        b["lat_img"].copy_(latents_img, non_blocking=True)
        b["lat_attr"].copy_(latents_attr, non_blocking=True)
        if not plan.cfg:
            b["ehs"].copy_(prompt_embeds.reshape(plan.B * plan.L, -1), non_blocking=True)

    def run(self, plan: _Plan, steps: Optional[int] = None):
        """setup + `steps` replays of the step program on the current stream (asynchronous)."""
        n = plan.steps if steps is None else steps
        if n > plan.steps:
            raise ValueError("more steps than the plan's schedule")
        if plan.ctx is not None:
            plan.ctx.sample_loop(n)            # one C call: zero the counter, setup, n x step (graph replays)
            return
        plan.bufs["step"].zero_()
        plan.setup.run()
        if self.use_graph:
            for _ in range(n):
                plan.step.launch_graph()
        else:
            for _ in range(n):
                plan.step.run()

    def launches_per_call(self, plan: _Plan) -> int:
        return plan.setup.num_launches + plan.steps * plan.step.num_launches

    def _sample(self, mode, latents_img, latents_attr, prompt_embeds, num_inference_steps, guidance_scale,
                scheduler=None, negative_prompt_embeds=None):
        # `do_classifier_free_guidance`: `guidance_scale != 0` in models/pipeline.py:807 (forward / inverse rendering),
        # but `guidance_scale > 1` in the joint loop's pipeline (models/pipeline_new_d4p.py:808-809)
        g = 0.0 if guidance_scale is None else float(guidance_scale)
        cfg = (g > 1.0) if mode in ("joint", "cycle") else (g != 0.0)
        if cfg and negative_prompt_embeds is None:
            raise ValueError("guidance_scale != 0 needs negative_prompt_embeds (every shipped Uni-Renderer caller passes "
                             "guidance_scale=0, eval/test_real.py:548)")
        B, _, S, S2 = latents_img.shape
        if S != S2:
            raise ValueError("square latents only")
        L = prompt_embeds.shape[1]
        plan = self.plan(mode, B, S, L, num_inference_steps, scheduler, cfg=cfg)
        self.load_inputs(plan, latents_img, latents_attr, prompt_embeds, negative_prompt_embeds if cfg else None,
                         guidance_scale if cfg else 0.0)
        self.run(plan)
        # always COPIES: the plan's latent buffers are static state that the next call with the same plan overwrites
        # (a same-device fp32 `.to()` would hand the caller an alias of them)
        dev, dt = latents_img.device, latents_img.dtype
        img = plan.bufs["lat_img"].to(device=dev, dtype=dt, non_blocking=False, copy=True)
        attr = plan.bufs["lat_attr"].to(device=dev, dtype=dt, non_blocking=False, copy=True)
        return plan, img, attr

    @torch.no_grad()
    def joint_sample(self, latents_img, latents_attr, prompt_embeds, num_inference_steps: int = 50,
                     guidance_scale: float = 0.0, scheduler: Optional[str] = None, negative_prompt_embeds=None):
        """Both streams noisy, same timestep (pipeline_new_d4p.py:1391-1453).  Returns (latents_img, latents_attr)."""
        _, img, attr = self._sample("joint", latents_img, latents_attr, prompt_embeds, num_inference_steps,
                                    guidance_scale, scheduler, negative_prompt_embeds)
        return img, attr

    @torch.no_grad()
    def forward_render(self, latents_img, attr_latents, prompt_embeds, num_inference_steps: int = 50,
                       guidance_scale: float = 0.0, scheduler: Optional[str] = None, negative_prompt_embeds=None):
        """attributes -> RGB: t_attr = 0, t_img: T -> 0 (pipeline.py:1455,1586-1653).  Returns latents_img."""
        return self._sample("forward", latents_img, attr_latents, prompt_embeds, num_inference_steps,
                            guidance_scale, scheduler, negative_prompt_embeds)[1]

    @torch.no_grad()
    def inverse_render(self, image_latents, latents_attr, prompt_embeds, num_inference_steps: int = 50,
                       guidance_scale: float = 0.0, scheduler: Optional[str] = None, negative_prompt_embeds=None):
        """RGB -> attributes: t_img = 0, t_attr: T -> 0 (pipeline.py:2475,2627-2733).  Returns the 24 attribute
        channels (the clean mask group is sliced off like pipeline.py:2691)."""
        return self._sample("inverse", image_latents, latents_attr, prompt_embeds, num_inference_steps,
                            guidance_scale, scheduler, negative_prompt_embeds)[2][:, MASK_CHANNELS:]

    @torch.no_grad()
    def cycle_sample(self, latents_img, latents_attr, prompt_embeds, num_inference_steps: int = 50,
                     guidance_scale: float = 0.0):
        """Joint step followed by the cycle-consistency pass of train/train.py:1388-1413 (attribute encoder on
        cat(mask, attr_pred) at t_attr = 0 -> RGB stream); the RGB update uses the second pass' prediction."""
        _, img, attr = self._sample("cycle", latents_img, latents_attr, prompt_embeds, num_inference_steps,
                                    guidance_scale)
        return img, attr


def _rows(a: Act, b0: int, b1: int) -> Act:
    """Samples [b0, b1) of an activation (a zero-copy row slice: the rows of a sample are contiguous)."""
    hw = a.H * a.W
    return Act(a.t[b0 * hw:b1 * hw], b1 - b0, a.H, a.W, a.C)


def _kv_rows(kv: Dict[str, torch.Tensor], b0: int, b1: int, L: int) -> Dict[str, torch.Tensor]:
    return {k: t[b0 * L:b1 * L] for k, t in kv.items()}


def all_gather_latents(x: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the sharded sampling loop: gather every rank's final latents [B_local, C, S, S] into
    [world * B_local, C, S, S] (NCCL all-gather over NVLink on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    world = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous batch split: [start, stop) of the samples rank `rank` denoises."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per
