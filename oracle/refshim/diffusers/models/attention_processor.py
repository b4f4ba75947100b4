"""Attention restated from the published diffusers-0.24 algorithm (AttnProcessor2_0 path, test-only)."""
import torch
import torch.nn.functional as F
from torch import nn


class AttnProcessor:
    pass


class AttnProcessor2_0:
    pass


class AttnAddedKVProcessor:
    pass


class AttnAddedKVProcessor2_0:
    pass


AttentionProcessor = AttnProcessor
ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor, AttnAddedKVProcessor2_0)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor, AttnProcessor2_0)


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, only_cross_attention=False, **unused):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        assert attention_mask is None, "no shipped caller passes a mask"
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = hidden_states.shape
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        d = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, n, self.heads * d).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))
