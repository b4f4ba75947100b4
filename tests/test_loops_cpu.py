"""The REAL recorders of the fused sampling loops (engine.StreamNet, pipeline.DualStreamSampler.plan: hoisting, lanes,
exchange epilogues, time-embedding tables, fused scheduler updates) executed on the CPU through
tests/cpu_ops_emulator.py and compared with the oracle's step-by-step loop -- host-side wiring of every mode without a
GPU -- and the sharded N > 1 path as a two-process gloo run.  Kernel parity: -m gpu tests."""
import os
import subprocess
import sys
from dataclasses import replace

import pytest
import torch

from oracle import uni_oracle as uo
from tests import cpu_ops_emulator as emu
from tests.sampler_probe import oracle_step

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(prediction_type="epsilon"):
    from uni_renderer_b200.engine import NetConfig
    base = uo.TINY
    cfgs_o = (replace(base), replace(base, in_channels=28), replace(base, out_channels=28))
    sds = [uo.random_state_dict(k, c, s) for k, c, s in zip(("unet", "attr_enc", "attr_dec"), cfgs_o, (11, 12, 13))]
    nb = NetConfig(block_out_channels=base.block_out_channels, num_heads=base.num_heads,
                   cross_attention_dim=base.cross_attention_dim, norm_num_groups=base.norm_num_groups)
    cfgs = (replace(nb), replace(nb, in_channels=28), replace(nb, out_channels=28))
    return emu.cpu_sampler(sds, cfgs, prediction_type), sds, cfgs_o


def _inputs(B, S, D, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 4, S, S, generator=g), torch.randn(B, 28, S, S, generator=g),
            torch.randn(B, 7, D, generator=g).half())


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("mode,scheduler,steps", [("joint", "ddim", 2), ("forward", "ddim", 2), ("inverse", "ddim", 2),
                                                   ("cycle", "ddim", 2), ("joint", "unipc", 3), ("inverse", "unipc", 3)])
def test_recorded_loop_matches_oracle(monkeypatch, mode, scheduler, steps):
    emu.install(monkeypatch)
    sampler, sds, cfgs = _setup()
    B, S, total = 2, 8, 10
    x_img, x_attr, ehs = _inputs(B, S, cfgs[0].cross_attention_dim)
    plan = sampler.plan(mode, B, S, ehs.shape[1], total, scheduler)
    sampler.load_inputs(plan, x_img, x_attr, ehs)
    sampler.run(plan, steps=steps)
    mk = (lambda: uo.DDIM()) if scheduler == "ddim" else (lambda: uo.UniPC())
    sched, sched_a = mk(), mk()
    ts = sched.set_timesteps(total)
    sched_a.set_timesteps(total)
    ri, ra = x_img, x_attr
    for i in range(steps):
        ri, ra = oracle_step(mode, sds, cfgs, sched, ts[i], ri, ra, ehs.float(), sched_a)
    got_i, got_a = plan.bufs["lat_img"], plan.bufs["lat_attr"]
    assert int(plan.bufs["step"].item()) == steps
    assert torch.equal(got_a[:, :4], x_attr[:, :4])            # the clean mask group is never updated
    if mode != "inverse":
        assert _rel(got_i, ri) < 5e-3, _rel(got_i, ri)
    else:
        assert torch.equal(got_i, x_img)                       # the RGB latent is an input of inverse rendering
    if mode != "forward":
        assert _rel(got_a, ra) < 5e-3, _rel(got_a, ra)
    else:
        assert torch.equal(got_a, x_attr)


@pytest.mark.parametrize("mode", ["forward", "inverse"])
def test_option_paths_agree_with_the_default(monkeypatch, mode):
    """A/B knobs of the recorder must not change the result: batch halves on two lanes (split_batch: row-sliced
    activations, K/V, time-embedding rows and latent views) and time embeddings recomputed inside every step
    (UNIB200_TEMB_TABLE=0) instead of tabulated for the whole loop."""
    emu.install(monkeypatch)
    _, sds, cfgs_o = _setup()
    from uni_renderer_b200.engine import NetConfig
    nb = NetConfig(block_out_channels=uo.TINY.block_out_channels, num_heads=uo.TINY.num_heads,
                   cross_attention_dim=uo.TINY.cross_attention_dim, norm_num_groups=uo.TINY.norm_num_groups)
    cfgs = (replace(nb), replace(nb, in_channels=28), replace(nb, out_channels=28))
    x_img, x_attr, ehs = _inputs(2, 8, nb.cross_attention_dim)
    outs = []
    for kw in ({}, {"split_batch": True}, {"temb_table": False}):
        s = emu.cpu_sampler(sds, cfgs, **kw)
        plan = s.plan(mode, 2, 8, ehs.shape[1], 10)
        s.load_inputs(plan, x_img, x_attr, ehs)
        s.run(plan, steps=2)
        outs.append((plan.bufs["lat_img"].clone(), plan.bufs["lat_attr"].clone()))
    for gi, ga in outs[1:]:
        assert _rel(gi, outs[0][0]) < 1e-3 and _rel(ga, outs[0][1]) < 1e-3


def test_plans_are_cached_and_modes_are_validated(monkeypatch):
    emu.install(monkeypatch)
    sampler, _, cfgs = _setup()
    p1 = sampler.plan("joint", 1, 8, 7, 4)
    assert sampler.plan("joint", 1, 8, 7, 4) is p1 and sampler.plan("joint", 1, 8, 7, 5) is not p1
    with pytest.raises(ValueError):
        sampler.plan("sideways", 1, 8, 7, 4)
    with pytest.raises(NotImplementedError):
        sampler.plan("cycle", 1, 8, 7, 4, "unipc")
    with pytest.raises(ValueError):
        sampler.load_inputs(p1, torch.zeros(2, 4, 8, 8), torch.zeros(1, 28, 8, 8), torch.zeros(1, 7, 48))
    with pytest.raises(ValueError):            # guidance needs the negative embeddings
        sampler.joint_sample(torch.zeros(1, 4, 8, 8), torch.zeros(1, 28, 8, 8), torch.zeros(1, 7, 48).half(), 4,
                             guidance_scale=7.5)
    with pytest.raises(NotImplementedError):
        sampler.plan("cycle", 1, 8, 7, 4, cfg=True)
    # step-invariant work is hoisted: forward rendering's step program is shorter than the joint one
    pj, pf = sampler.plan("joint", 1, 8, 7, 4), sampler.plan("forward", 1, 8, 7, 4)
    assert pf.step.num_launches < 0.7 * pj.step.num_launches and pf.setup.num_launches > pj.setup.num_launches


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
import pytest
from tests import cpu_ops_emulator as emu
from tests.test_loops_cpu import _setup, _inputs
from uni_renderer_b200.pipeline import all_gather_latents, shard_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mp = pytest.MonkeyPatch()
emu.install(mp)
sampler, sds, cfgs = _setup()
GB, S = 4, 8
x_img, x_attr, ehs = _inputs(GB, S, cfgs[0].cross_attention_dim)          # the same global batch on every rank
lo, hi = shard_batch(GB, rank, world)
img, attr = sampler.joint_sample(x_img[lo:hi], x_attr[lo:hi], ehs[lo:hi], num_inference_steps=2)
out = all_gather_latents(torch.cat([img, attr], 1))                       # the one collective of the sharded loop
if rank == 0:
    full_i, full_a = sampler.joint_sample(x_img, x_attr, ehs, num_inference_steps=2)     # unsharded run of the same batch
    ref = torch.cat([full_i, full_a], 1)
    assert out.shape == ref.shape
    # images are independent; the two runs differ only by fp32 summation order inside the emulator's matmuls (the
    # batch changes the BLAS blocking), which flips a few fp16 roundings: measured rel_l2 2.8e-4
    rel = ((out - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, rel
    assert torch.equal(out[:, 4:8], torch.cat([x_attr[:, :4]], 0))                       # gather order = batch order
dist.barrier()
dist.destroy_process_group()
mp.undo()
print("OK", rank)
"""


def test_sharded_loop_world2_gloo(tmp_path):
    """Two ranks denoise the two halves of a batch and all-gather the final latents; the result equals the
    single-process run of the whole batch (no cross-sample op anywhere in the loop)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("OK") == 2


@pytest.mark.parametrize("name", ["tiny_step_vec_t.pt", "tiny_step_scalar_t.pt"])
def test_module_forwards_on_emulator_match_reference_golden(monkeypatch, golden_dir, name):
    """The three drop-in nn.Modules called in the reference's 3-call sequence (train/train.py:1324-1354), their recorded
    programs executed by the emulator, against the tensors recorded from the REFERENCE's own model files
    (tests/golden, oracle/make_golden.py): the module-level host wiring (timestep forms, residual ingestion, skip /
    tap order, return structures) is pinned to the reference in the CPU suite too."""
    from tests import gpu_model_probe as gp
    from uni_renderer_b200 import models as M
    from uni_renderer_b200.engine import StreamNet, Workspace
    emu.install(monkeypatch)

    def cpu_finalize(self, device=None):            # test only: the product's finalize() refuses non-CUDA devices
        if self._net is None:
            self._net = StreamNet(self._kind, self.net_cfg, dict(self.state_dict()), "cpu")
            self._ws = Workspace("cpu")
        return self._net
    monkeypatch.setattr(M._NetModule, "finalize", cpu_finalize)
    gold = torch.load(os.path.join(golden_dir, name), weights_only=False)
    gc = gold["config"]
    (unet, enc, dec), sds, _ = gp.build_modules(gc, device="cpu")
    for sd, dg in zip(sds, gold["weight_digest"]):
        if abs(float(sum(v.double().sum() for v in sd.values())) - dg["sum"]) > 1e-6 * max(1.0, abs(dg["abs"])):
            pytest.skip("torch CPU RNG stream differs from the one that generated the fixtures")
    B = gc["B"]
    ti, ta = (gc["t_img"], gc["t_attr"]) if gc["scalar_t"] else (torch.full((B,), gc["t_img"]), torch.full((B,), gc["t_attr"]))
    x_img, x_attr, ehs = gold["x_img"], gold["x_attr"], gold["ehs"]
    d, m, raw_a, raw_a_mid = enc(x_img, ta, encoder_hidden_states=ehs, controlnet_cond=x_attr, return_dict=False)
    img, raw_u, raw_u_mid, taps = unet(x_img, ti, encoder_hidden_states=ehs, down_block_additional_residuals=d,
                                       mid_block_additional_residual=m, return_dict=False)
    attr = dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=ta, encoder_hidden_states=ehs,
               down_block_additional_residuals=raw_u, mid_block_additional_residual=raw_u_mid, return_dict=False)
    assert len(d) == 12 and len(raw_a) == 12 and len(raw_u) == 12 and len(taps) == 13
    checks = [("unet_sample", img, gold["unet_sample"]), ("dec_sample", attr, gold["dec_sample"]),
              ("enc_mid", m, gold["enc_mid"]), ("enc_raw_mid", raw_a_mid, gold["enc_raw_mid"]),
              ("unet_raw_mid", raw_u_mid, gold["unet_raw_mid"])]
    checks += [(f"enc_down{i}", d[i], gold["enc_down"][i]) for i in range(12)]
    checks += [(f"enc_raw{i}", raw_a[i], gold["enc_raw_down"][i]) for i in range(12)]
    checks += [(f"unet_raw{i}", raw_u[i], gold["unet_raw_down"][i]) for i in range(12)]
    checks += [(f"unet_tap{i}", taps[i], gold["unet_up_taps"][i]) for i in range(13)]
    for what, got, ref in checks:
        assert tuple(got.shape) == tuple(ref.shape), what
        assert _rel(got.float(), ref) < 3e-3, (what, _rel(got.float(), ref))
    plain = unet(x_img, ti, encoder_hidden_states=ehs).sample
    assert _rel(plain.float(), gold["unet_sample_plain"]) < 3e-3
    # conditioning_scale is folded into the 13 zero-convs (models/controlnet.py:1773-1775)
    d2, m2, _, _ = enc(x_img, ta, encoder_hidden_states=ehs, controlnet_cond=x_attr, conditioning_scale=0.5,
                       return_dict=False)
    assert _rel(m2.float(), 0.5 * gold["enc_mid"]) < 3e-3 and _rel(d2[3].float(), 0.5 * gold["enc_down"][3]) < 3e-3


def test_render_pipeline_on_emulator_matches_oracle_chain(monkeypatch):
    """Image-to-image inverse rendering (RenderPipeline: 2 batched encodes -> loop -> 5 batched decodes) on the CPU
    emulator against vae_oracle.encode -> uni_oracle steps -> vae_oracle.decode with the same seeded generator: pins
    the noise order (posterior noise of image then masks, then the six attribute latents; models/pipeline.py:2112-2188),
    the 28-channel layout (:1583) and the decode order (:2335-2349) in the CPU suite."""
    from oracle import vae_oracle as vo
    from uni_renderer_b200 import vae as V
    from uni_renderer_b200.engine import Workspace
    from uni_renderer_b200.render import RenderPipeline
    emu.install(monkeypatch)
    sampler, sds, cfgs = _setup()
    vsd = vo.random_state_dict(vo.TINY_VAE, 5)
    vae = V.AutoencoderKL(block_out_channels=(32, 64, 64), down_block_types=(V._DOWN,) * 3, up_block_types=(V._UP,) * 3,
                          layers_per_block=2, norm_num_groups=8)
    vae.load_state_dict(vsd)
    vae.use_graph = False

    def cpu_finalize(self, device=None):            # test only: bypass the CUDA gate of the product
        if self._net is None:
            net = object.__new__(V.VaeNet)
            net.cfg, net.device, net.w = self.vae_cfg, torch.device("cpu"), {}
            net._pack(V.convert_deprecated_attention_keys(dict(self.state_dict())))
            self._net, self._ws = net, Workspace("cpu")
        return self._net
    monkeypatch.setattr(V.AutoencoderKL, "finalize", cpu_finalize)
    rp = RenderPipeline(sampler, vae)
    B, S, steps = 1, 32, 2
    h = S // 4
    g = torch.Generator().manual_seed(31)
    image, masks = (torch.tanh(torch.randn(B, 3, S, S, generator=g)) for _ in range(2))
    ehs = torch.randn(B, 7, cfgs[0].cross_attention_dim, generator=g).half()
    gen = torch.Generator().manual_seed(77)
    out = rp.inverse_rendering(image, masks, ehs, num_inference_steps=steps, generator=gen, posterior_generator=gen)
    assert len(out) == 6 and out[0].shape == (B, 4, h, h) and all(o.shape == (B, 3, S, S) for o in out[1:])
    gen.manual_seed(77)
    n_img, n_msk = (torch.randn(B, 4, h, h, generator=gen) for _ in range(2))
    lat = [torch.randn(B, 4, h, h, generator=gen) for _ in range(6)]
    sf = vo.TINY_VAE.scaling_factor
    with torch.no_grad():
        l_img = vo.sample_posterior(vo.encode_moments(vsd, vo.TINY_VAE, image), n_img) * sf
        l_msk = vo.sample_posterior(vo.encode_moments(vsd, vo.TINY_VAE, masks), n_msk) * sf
        sched, sched_a = uo.DDIM(), uo.DDIM()
        ts = sched.set_timesteps(steps)
        sched_a.set_timesteps(steps)
        xi, xa = l_img, torch.cat([l_msk] + lat, 1)
        for i in range(steps):
            xi, xa = oracle_step("inverse", sds, cfgs, sched, ts[i], xi, xa, ehs.float(), sched_a)
        assert _rel(out[0], xa[:, 4:8]) < 5e-3
        for i in range(5):
            ref = vo.decode(vsd, vo.TINY_VAE, xa[:, 8 + 4 * i:12 + 4 * i] / sf)
            assert _rel(out[1 + i], ref) < 6e-3, (i, _rel(out[1 + i], ref))
    # prompt embeddings come from the cache when none are passed
    with pytest.raises(ValueError):
        rp.inverse_rendering(image, masks, None, num_inference_steps=steps)
    # guidance passes through to the CFG plan of the loop
    neg = torch.randn(1, 7, cfgs[0].cross_attention_dim, generator=g).half()
    gen.manual_seed(77)
    guided = rp.inverse_rendering(image, masks, ehs, num_inference_steps=steps, generator=gen, guidance_scale=2.0,
                                  negative_prompt_embeds=neg, posterior_generator=gen)
    assert _rel(guided[0], out[0]) > 1e-3 and _rel(guided[1], out[1]) < 0.5     # material group is guided (:2263)


@pytest.mark.parametrize("mode,scheduler,steps", [("joint", "ddim", 2), ("forward", "ddim", 2), ("inverse", "ddim", 2),
                                                   ("joint", "unipc", 3), ("inverse", "unipc", 3)])
def test_cfg_plans_match_the_reference_rules(monkeypatch, mode, scheduler, steps):
    """Classifier-free guidance (`guidance_scale != 0`, models/pipeline.py:807): doubled batch with
    cat([negative, positive]) embeddings and the combination rule of each loop (tests/cfg_reference.py), through the
    public entry points on the emulator."""
    from tests.cfg_reference import cfg_step
    emu.install(monkeypatch)
    sampler, sds, cfgs = _setup()
    B, S, total, g = 2, 8, 10, 3.0
    x_img, x_attr, ehs = _inputs(B, S, cfgs[0].cross_attention_dim)
    neg = torch.randn(1, ehs.shape[1], ehs.shape[2], generator=torch.Generator().manual_seed(5)).half()
    plan = sampler.plan(mode, B, S, ehs.shape[1], total, scheduler, cfg=True)
    assert plan.cfg and plan.net_batch == 2 * B and plan.bufs["lat_img"].shape[0] == B
    sampler.load_inputs(plan, x_img, x_attr, ehs, neg, g)
    sampler.run(plan, steps=steps)
    mk = (lambda: uo.DDIM()) if scheduler == "ddim" else (lambda: uo.UniPC())
    sched, sched_a = mk(), mk()
    ts = sched.set_timesteps(total)
    sched_a.set_timesteps(total)
    ri, ra = x_img, x_attr
    for i in range(steps):
        ri, ra = cfg_step(mode, sds, cfgs, sched, ts[i], ri, ra, ehs.float(), neg.float(), g, sched_a)
    got_i, got_a = plan.bufs["lat_img"], plan.bufs["lat_attr"]
    assert torch.equal(got_a[:, :4], x_attr[:, :4])
    if mode != "inverse":
        assert _rel(got_i, ri) < 5e-3, _rel(got_i, ri)
    if mode != "forward":
        assert _rel(got_a, ra) < 5e-3, _rel(got_a, ra)
    # guidance actually changes the result (the plan without guidance gives something else)
    p0 = sampler.plan(mode, B, S, ehs.shape[1], total, scheduler)
    sampler.load_inputs(p0, x_img, x_attr, ehs)
    sampler.run(p0, steps=steps)
    tgt = "lat_attr" if mode == "inverse" else "lat_img"
    assert _rel(p0.bufs[tgt], plan.bufs[tgt]) > 1e-3
    with pytest.raises(ValueError):
        sampler.load_inputs(plan, x_img, x_attr, ehs)              # a CFG plan needs the negative embeddings
    with pytest.raises(ValueError):
        sampler.load_inputs(p0, x_img, x_attr, ehs, neg, g)


@pytest.mark.parametrize("mode", ["joint", "forward"])
def test_groupnorm_statistics_fused_into_gemm_epilogues(monkeypatch, mode):
    """North-star item "GroupNorm fused into the epilogue": with the fusion forced on at every size the kernels allow,
    every GEMM that produces a GroupNorm input also writes (row block, micro-group) statistics, and the GroupNorm of a
    single source AND of a two-source concat (decoder: h | exchanged skip, other group size) normalises with the sums of
    those tables -- the emulator computes the statistics ONLY from the tables, so a table handed to the wrong norm, a
    wrong granularity or a missing producer shows up as a mismatch with the oracle."""
    emu.install(monkeypatch)
    sampler, sds, cfgs = _setup()
    for net in (sampler.unet, sampler.enc, sampler.dec):
        assert net.gn_gran == 4                          # gcd of the tiny config's group sizes (32 / 8, 64 / 8, 128 / 8)
        net.gn_force = True
    B, S, total = 2, 16, 10                              # 16x16, 8x8 fuse (HW % 32 == 0); 4x4 and 2x2 fall back
    seen = {"fused": 0, "plain": 0, "producers": 0}
    from uni_renderer_b200 import ops
    real_gn, real_gemm = ops.groupnorm, ops.conv_gemm

    def spy_gn(*a, **k):
        seen["fused" if k.get("parts") is not None else "plain"] += 1
        if k.get("parts") is not None and a[3] is not None:
            seen["two_source"] = seen.get("two_source", 0) + 1
        return real_gn(*a, **k)

    def spy_gemm(*a, **k):
        seen["producers"] += k.get("gn") is not None
        return real_gemm(*a, **k)
    monkeypatch.setattr(ops, "groupnorm", spy_gn)
    monkeypatch.setattr(ops, "conv_gemm", spy_gemm)
    x_img, x_attr, ehs = _inputs(B, S, cfgs[0].cross_attention_dim)
    plan = sampler.plan(mode, B, S, ehs.shape[1], total, "ddim")
    sampler.load_inputs(plan, x_img, x_attr, ehs)
    sampler.run(plan, steps=2)
    assert seen["fused"] > 20 and seen["plain"] > 0 and seen["producers"] >= seen["fused"]
    if mode == "joint":
        assert seen.get("two_source", 0) >= 6            # decoder concat norms at 16x16 and 8x8, both streams
    sched, sched_a = uo.DDIM(), uo.DDIM()
    ts = sched.set_timesteps(total)
    sched_a.set_timesteps(total)
    ri, ra = x_img, x_attr
    for i in range(2):
        ri, ra = oracle_step(mode, sds, cfgs, sched, ts[i], ri, ra, ehs.float(), sched_a)
    assert _rel(plan.bufs["lat_img"], ri) < 5e-3
    if mode == "joint":
        assert _rel(plan.bufs["lat_attr"], ra) < 5e-3


def test_upres_decoder_blocks_add_the_extra_residual_per_layer(monkeypatch):
    """UpResBlock2D / CrossAttnUpResBlock2D (`hidden_states += up_additional_states` after every decoder
    layer, models/unet_2d_blocks.py:2408,2814) through the drop-in AttributeDecoderModel with its class-default
    up_block_types, on the emulator against the oracle.  The add rides through the layer's last GEMM (epilogue residual,
    or an identity K segment where that input is taken) -- no kernel of its own.  With the plain up blocks the argument
    is ignored with a warning, exactly like the reference's live forward (models/controlnet.py:2464-2510)."""
    from tests import gpu_model_probe as gp
    from uni_renderer_b200 import models as M
    from uni_renderer_b200.engine import StreamNet, Workspace
    emu.install(monkeypatch)

    def cpu_finalize(self, device=None):
        if self._net is None:
            self._net = StreamNet(self._kind, self.net_cfg, dict(self.state_dict()), "cpu")
            self._ws = Workspace("cpu")
        return self._net
    monkeypatch.setattr(M._NetModule, "finalize", cpu_finalize)
    gc = dict(block_out_channels=uo.TINY.block_out_channels, num_heads=uo.TINY.num_heads,
              cross_attention_dim=uo.TINY.cross_attention_dim, norm_num_groups=uo.TINY.norm_num_groups, seeds=(11, 12, 13))
    (unet, enc, dec_plain), sds, cfgs = gp.build_modules(gc, device="cpu")
    kw = dict(block_out_channels=tuple(gc["block_out_channels"]), attention_head_dim=gc["num_heads"],
              cross_attention_dim=gc["cross_attention_dim"], norm_num_groups=gc["norm_num_groups"])
    dec = M.AttributeDecoderModel(out_channels=28, _init_weights=False, **kw)            # UpRes defaults
    dec.load_state_dict(sds[2])
    B, S = 2, 8
    x_img, x_attr, ehs = _inputs(B, S, cfgs[0].cross_attention_dim)
    ehs = ehs.float()
    t = 501
    d, m, raw_a, raw_a_mid = enc(x_img, t, encoder_hidden_states=ehs, controlnet_cond=x_attr, return_dict=False)
    _, raw_u, raw_u_mid, taps = unet(x_img, t, encoder_hidden_states=ehs, down_block_additional_residuals=d,
                                     mid_block_additional_residual=m, return_dict=False)
    g = torch.Generator().manual_seed(9)
    ups = [0.5 * torch.randn(tp.shape, generator=g) for tp in taps[1:]]                  # one per decoder layer
    got = dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=t, encoder_hidden_states=ehs,
              down_block_additional_residuals=raw_u, up_block_additional_residuals=ups,
              mid_block_additional_residual=raw_u_mid, return_dict=False)
    f = lambda ts: [x.float() for x in ts]       # noqa: E731
    with torch.no_grad():
        ref = uo.attr_decoder_forward(sds[2], cfgs[2], raw_a_mid.float(), f(raw_a), t, ehs, f(raw_u), raw_u_mid.float(),
                                      up_block_additional_residuals=[u.half().float() for u in ups])
        ref_plain = uo.attr_decoder_forward(sds[2], cfgs[2], raw_a_mid.float(), f(raw_a), t, ehs, f(raw_u),
                                            raw_u_mid.float())
    assert _rel(got.float(), ref) < 3e-3 and _rel(ref_plain, ref) > 0.05
    with pytest.raises(ValueError):
        dec(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=t, encoder_hidden_states=ehs,
            down_block_additional_residuals=raw_u, mid_block_additional_residual=raw_u_mid, return_dict=False)
    with pytest.warns(UserWarning):
        ign = dec_plain(sample=raw_a_mid, down_block_res_samples=raw_a, timestep=t, encoder_hidden_states=ehs,
                        down_block_additional_residuals=raw_u, up_block_additional_residuals=ups,
                        mid_block_additional_residual=raw_u_mid, return_dict=False)
    assert _rel(ign.float(), ref_plain) < 3e-3
