"""Attention + AttnProcessor / AttnProcessor2_0 as in diffusers 0.24.0 (restated; only the options
the reference's hot path uses: no added_kv, no group/spatial norm, no norm_cross, no LoRA)."""
from typing import Union

import torch
import torch.nn.functional as F
from torch import nn


class AttnProcessor:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale=1.0):
        batch_size = hidden_states.shape[0]
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        h = attn.heads
        d = key.shape[-1] // h

        def split(t):
            return t.view(batch_size, -1, h, d).transpose(1, 2).reshape(batch_size * h, -1, d)

        q, k, v = split(query), split(key), split(value)
        scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype,
                                           device=q.device), q, k.transpose(-1, -2), beta=0,
                               alpha=attn.scale)
        probs = scores.softmax(dim=-1).to(v.dtype)
        out = torch.bmm(probs, v)
        out = out.view(batch_size, h, -1, d).transpose(1, 2).reshape(batch_size, -1, h * d)
        out = attn.to_out[0](out)
        out = attn.to_out[1](out)
        return out / attn.rescale_output_factor


class AttnProcessor2_0:
    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale=1.0):
        batch_size = (hidden_states if encoder_hidden_states is None else encoder_hidden_states).shape[0]
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        inner_dim = key.shape[-1]
        head_dim = inner_dim // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        hidden_states = F.scaled_dot_product_attention(query, key, value, attn_mask=attention_mask,
                                                       dropout_p=0.0, is_causal=False)
        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim)
        hidden_states = hidden_states.to(query.dtype)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        return hidden_states / attn.rescale_output_factor


AttentionProcessor = Union[AttnProcessor, AttnProcessor2_0]


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0,
                 bias=False, upcast_attention=False, upcast_softmax=False, cross_attention_norm=None,
                 cross_attention_norm_num_groups=32, added_kv_proj_dim=None, norm_num_groups=None,
                 spatial_norm_dim=None, out_bias=True, scale_qk=True, only_cross_attention=False,
                 eps=1e-5, rescale_output_factor=1.0, residual_connection=False,
                 _from_deprecated_attn_block=False, processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.dropout = dropout
        self.scale_qk = scale_qk
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.only_cross_attention = only_cross_attention
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias),
                                     nn.Dropout(dropout)])
        if processor is None:
            processor = AttnProcessor2_0() if hasattr(F, "scaled_dot_product_attention") and scale_qk \
                else AttnProcessor()
        self.set_processor(processor)

    def set_processor(self, processor, _remove_lora=False):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None,
                **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class AttnAddedKVProcessor:  # names imported by the reference's unet_2d_mix.py; unused for SD-1.5
    pass


ADDED_KV_ATTENTION_PROCESSORS = (AttnAddedKVProcessor,)
CROSS_ATTENTION_PROCESSORS = (AttnProcessor, AttnProcessor2_0)
