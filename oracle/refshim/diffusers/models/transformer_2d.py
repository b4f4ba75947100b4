"""Transformer2DModel / BasicTransformerBlock / GEGLU restated from the published diffusers-0.24 algorithm
for use_linear_projection=False, num_layers=1, norm_type=layer_norm (test-only)."""
import torch
import torch.nn.functional as F
from torch import nn

from .attention_processor import Attention


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale=1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim)])

    def forward(self, x, scale=1.0):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim, only_cross_attention=False, upcast_attention=False):
        super().__init__()
        assert not only_cross_attention
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=True)
        self.attn1 = Attention(query_dim=dim, heads=heads, dim_head=dim_head, bias=False)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=True)
        self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=heads, dim_head=dim_head,
                               bias=False)
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=True)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        hidden_states = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states) + hidden_states
        hidden_states = self.ff(self.norm3(hidden_states)) + hidden_states
        return hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None,
                 num_layers=1, dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 attention_type="default", **unused):
        super().__init__()
        assert not use_linear_projection and attention_type == "default"
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention)
            for _ in range(num_layers)])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None, return_dict=True):
        assert attention_mask is None and encoder_attention_mask is None
        b, _, h, w = hidden_states.shape
        residual = hidden_states
        hidden_states = self.proj_in(self.norm(hidden_states))
        inner = hidden_states.shape[1]
        hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(b, h * w, inner)
        for blk in self.transformer_blocks:
            hidden_states = blk(hidden_states, encoder_hidden_states=encoder_hidden_states)
        hidden_states = hidden_states.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
        output = self.proj_out(hidden_states) + residual
        return (output,)
