"""Prompt-embedding cache against the reference's own text-encoder dependency: transformers'
CLIPTextModel (random-init, small config) executed directly, i.e. the oracle here is the real library."""
import pytest
import torch

transformers = pytest.importorskip("transformers")

from uni_renderer_b200.text import PromptEmbedCache  # noqa: E402


class _Tok:
    """Stand-in for CLIPTokenizer (its vocabulary files are not available offline): byte-level ids, BOS/EOS, padded to
    model_max_length -- the cache only relies on the call signature and `.input_ids` / `.attention_mask`."""
    model_max_length = 16

    class _Out:
        pass

    def __call__(self, prompt, padding=None, max_length=None, truncation=None, return_tensors=None):
        prompts = [prompt] if isinstance(prompt, str) else prompt
        ids, mask = [], []
        for p in prompts:
            t = [1] + [3 + (b % 90) for b in p.encode()][:max_length - 2] + [2]
            mask.append([1] * len(t) + [0] * (max_length - len(t)))
            ids.append(t + [2] * (max_length - len(t)))
        o = self._Out()
        o.input_ids, o.attention_mask = torch.tensor(ids), torch.tensor(mask)
        return o


def _encoder(calls):
    cfg = transformers.CLIPTextConfig(vocab_size=100, hidden_size=32, intermediate_size=64, num_hidden_layers=3,
                                      num_attention_heads=4, max_position_embeddings=16, bos_token_id=1, eos_token_id=2)
    torch.manual_seed(0)
    enc = transformers.CLIPTextModel(cfg).eval()
    orig = enc.forward

    def counted(*a, **k):
        calls.append(1)
        return orig(*a, **k)
    enc.forward = counted
    return enc


def test_cache_matches_direct_text_model_and_runs_it_once():
    calls = []
    enc, tok = _encoder(calls), _Tok()
    cache = PromptEmbedCache(tok, enc, device="cpu", dtype=torch.float32)
    e1 = cache.encode(" ")
    with torch.no_grad():
        ref = enc(tok(" ", max_length=16).input_ids)[0]
    n = len(calls)
    torch.testing.assert_close(e1, ref)
    for _ in range(5):
        assert torch.equal(cache.encode(" "), e1)
    assert len(calls) == n and cache.misses == 1 and cache.hits == 5
    # batch of prompts + num_images_per_prompt (models/pipeline.py:366-369)
    e = cache.encode([" ", "a b"], num_images_per_prompt=2)
    assert e.shape == (4, 16, 32) and torch.equal(e[0], e1[0]) and torch.equal(e[1], e1[0])
    with torch.no_grad():
        ref2 = enc(tok("a b", max_length=16).input_ids)[0]
    torch.testing.assert_close(e[2], ref2[0])
    assert cache.misses == 2


def test_clip_skip_reapplies_final_layer_norm():
    calls = []
    enc, tok = _encoder(calls), _Tok()
    cache = PromptEmbedCache(tok, enc, device="cpu", dtype=torch.float32)
    e = cache.encode(" ", clip_skip=1)
    with torch.no_grad():
        out = enc(tok(" ", max_length=16).input_ids, output_hidden_states=True)
        ref = enc.text_model.final_layer_norm(out[-1][-2])
    torch.testing.assert_close(e, ref)
    assert not torch.allclose(e, cache.encode(" "))        # a different cache entry than clip_skip=None
    assert cache.misses == 2


def test_lru_bound_and_fp16_default():
    enc, tok = _encoder([]), _Tok()
    cache = PromptEmbedCache(tok, enc, device="cpu", max_entries=2)
    for p in ("a", "b", "c"):
        assert cache.encode(p).dtype == torch.float16
    assert len(cache._cache) == 2 and ("a", None) not in cache._cache
