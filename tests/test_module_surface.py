"""CPU checks of the drop-in module surface: constructor / config / state-dict layout /
from_unet / diffusers-style persistence.  No compute runs here -- forward() without CUDA must raise."""
import json
import os
from dataclasses import replace

import pytest
import torch

from oracle import uni_oracle as uo
from uni_renderer_b200 import models as M

TINY = dict(block_out_channels=(32, 64, 128, 128), attention_head_dim=4, cross_attention_dim=48, norm_num_groups=8)


def _tiny_unet(**kw):
    return M.UNet2DConditionModel(in_channels=4, out_channels=4, **dict(TINY, **kw))


def test_state_dict_layout_equals_oracle_layout():
    """The oracle's state dicts load strict=True into the reference's own classes (oracle/make_golden.py), so equal
    key sets and shapes here mean our modules ingest reference checkpoints unchanged."""
    cfgs = (replace(uo.TINY), replace(uo.TINY, in_channels=28), replace(uo.TINY, out_channels=28))
    mods = (_tiny_unet(), M.AttributeEncoderModel(in_channels=28, **TINY),
            M.AttributeDecoderModel(out_channels=28, up_block_types=M._SD_UP, **TINY))
    for kind, cfg, m in zip(("unet", "attr_enc", "attr_dec"), cfgs, mods):
        ref = uo.random_state_dict(kind, cfg, 3)
        own = m.state_dict()
        assert set(own) == set(ref), (kind, sorted(set(own) ^ set(ref))[:8])
        for k in ref:
            assert tuple(own[k].shape) == tuple(ref[k].shape), (kind, k)
        m.load_state_dict(ref, strict=True)
        with pytest.raises(RuntimeError):
            m.load_state_dict({k: v for k, v in list(ref.items())[1:]}, strict=True)


def test_config_is_mutable_attribute_dict():
    u = _tiny_unet()
    assert u.config.in_channels == 4 and u.config["block_out_channels"] == (32, 64, 128, 128)
    assert "cross_attention_dim" in u.config
    u.config["in_channels"] = 8                    # train/train.py:985,996 mutates the config in place
    assert u.config.in_channels == 8
    with pytest.raises(AttributeError):
        u.config.no_such_key
    assert u.dtype == torch.float16


def test_constructor_rejects_what_the_path_does_not_implement():
    with pytest.raises(ValueError):
        _tiny_unet(num_attention_heads=8)          # same restriction as the reference (controlnet.py:212)
    with pytest.raises(ValueError):
        _tiny_unet(use_linear_projection=True)
    with pytest.raises(ValueError):
        _tiny_unet(down_block_types=("DownBlock2D",) * 4)
    # the decoder's class-default UpRes up blocks (models/controlnet.py:1794-1797) are constructible: SURVEY row a8
    d = M.AttributeDecoderModel(out_channels=28, **TINY)
    assert d.config.up_block_types == M._SD_UPRES and d.net_cfg.up_res
    assert set(d.state_dict()) == set(M.AttributeDecoderModel(out_channels=28, up_block_types=M._SD_UP, **TINY).state_dict())
    with pytest.raises(ValueError):
        M.AttributeDecoderModel(out_channels=28, up_block_types=("UpBlock2D",) * 4, **TINY)


def test_from_unet_copies_the_shared_trunk_and_zeroes_the_exchange():
    u = _tiny_unet()
    enc = M.AttributeEncoderModel.from_unet(u)
    dec = M.AttributeDecoderModel.from_unet(u)
    su, se, sd = u.state_dict(), enc.state_dict(), dec.state_dict()
    for k, v in se.items():
        if k.startswith(("controlnet_",)):
            assert float(v.abs().max()) == 0.0, k                   # zero_module(), controlnet.py:1360-1415
        else:
            assert torch.equal(v, su[k]), k
    for k, v in sd.items():
        if k.startswith("control_"):
            assert float(v.abs().max()) == 0.0, k                   # controlnet.py:1988-2009
        else:
            assert torch.equal(v, su[k]), k
    assert dec.config.up_block_types == u.config.up_block_types     # from_unet takes the UNet's block types (:2115)
    assert len([k for k in se if k.startswith("controlnet_down_blocks.") and k.endswith(".weight")]) == 12


@pytest.mark.parametrize("safe", [True, False])
def test_save_pretrained_from_pretrained_roundtrip(tmp_path, safe):
    enc = M.AttributeEncoderModel(in_channels=28, **TINY)
    with torch.no_grad():
        for p in enc.parameters():
            p.copy_(torch.randn_like(p))
    d = tmp_path / "ckpt" / "controlnet"
    enc.save_pretrained(str(d), safe_serialization=safe)
    cfg = json.load(open(d / "config.json"))
    assert cfg["_class_name"] == "AttributeEncoderModel" and cfg["in_channels"] == 28
    assert os.path.isfile(d / ("diffusion_pytorch_model.safetensors" if safe else "diffusion_pytorch_model.bin"))
    back = M.AttributeEncoderModel.from_pretrained(str(tmp_path / "ckpt"), subfolder="controlnet")
    assert dict(back.config) == dict(enc.config)
    a, b = enc.state_dict(), back.state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_from_pretrained_ignores_foreign_config_keys_and_fails_on_missing_files(tmp_path):
    u = _tiny_unet()
    u.save_pretrained(str(tmp_path / "unet"))
    p = tmp_path / "unet" / "config.json"
    cfg = json.load(open(p))
    cfg["_name_or_path"] = "somewhere"
    cfg["a_key_of_a_newer_diffusers"] = 1
    json.dump(cfg, open(p, "w"))
    back = M.UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    assert back.config.block_out_channels == (32, 64, 128, 128)
    with pytest.raises(EnvironmentError):
        M.UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="nope")
    os.remove(tmp_path / "unet" / "diffusion_pytorch_model.safetensors")
    with pytest.raises(EnvironmentError):
        M.UNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet")


def test_forward_without_cuda_raises_instead_of_falling_back():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    u = _tiny_unet()
    x = torch.zeros(1, 4, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        u(x, 10, torch.zeros(1, 7, 48))


def test_timestep_forms():
    f = M._NetModule._timesteps
    assert f(5, 3, "cpu").tolist() == [5.0, 5.0, 5.0]
    assert f(torch.tensor(7), 2, "cpu").tolist() == [7.0, 7.0]
    assert f(torch.tensor([1, 2]), 2, "cpu").tolist() == [1.0, 2.0]
    with pytest.raises(ValueError):
        f(torch.tensor([1, 2, 3]), 2, "cpu")


def test_remaining_module_methods_raise_clear_errors_instead_of_attribute_errors():
    """models/controlnet.py:591-779: attn_processors / set_attn_processor / set_default_attn_processor /
    set_attention_slice / enable_freeu / disable_freeu exist on the reference modules; here the switches that would
    change the fused kernels raise NotImplementedError, the ones that restore the default are no-ops."""
    for m in (_tiny_unet(), M.AttributeEncoderModel(in_channels=28, **TINY),
              M.AttributeDecoderModel(out_channels=28, up_block_types=M._SD_UP, **TINY)):
        assert m.attn_processors == {}
        assert m.set_default_attn_processor() is None and m.disable_freeu() is None
        assert m.enable_xformers_memory_efficient_attention() is None
        with pytest.raises(NotImplementedError):
            m.set_attn_processor(object())
        with pytest.raises(NotImplementedError):
            m.set_attention_slice("auto")
        with pytest.raises(NotImplementedError):
            m.enable_freeu(s1=0.9, s2=0.2, b1=1.2, b2=1.4)
