"""CPU: the oracle's restatements of the diffusers-owned primitives ("restatement-pinned" — diffusers is
neither under /root/reference nor installed) against INDEPENDENT implementations that are in this image:

* `Attention` + `AttnProcessor2_0`  vs  `torch.nn.MultiheadAttention` (separate q / k / v projection weights, kdim / vdim
  for the CLIP cross-attention, no projection bias on q / k / v, bias on the output projection, scale d^-0.5),
* `Timesteps`  vs  transformers' tensor2tensor sinusoidal table (`Speech2TextSinusoidalPositionalEmbedding`: the same
  family with downscale_freq_shift = 1 and [sin | cos] order) and vs the closed form in float64 for the reference's
  own parameters (flip_sin_to_cos = True, freq_shift = 0: src/models/unet_3d_mix.py:153-155),
* GEGLU's activation  vs  the GELU paper's closed form  x/2 (1 + erf(x / sqrt 2))  (diffusers' default, not the tanh
  approximation), with [value | gate] halves.
"""
import math

import pytest
import torch

from oracle import unet3d_oracle as O


@pytest.mark.parametrize("cross_dim", [None, 24])
def test_attention_matches_torch_multihead_attention(cross_dim):
    torch.manual_seed(0)
    C, heads, B, Lq, Lkv = 32, 4, 2, 7, 5
    kv = cross_dim or C
    sd = {"a.to_q.weight": torch.randn(C, C) * 0.2, "a.to_k.weight": torch.randn(C, kv) * 0.2,
          "a.to_v.weight": torch.randn(C, kv) * 0.2, "a.to_out.0.weight": torch.randn(C, C) * 0.2,
          "a.to_out.0.bias": torch.randn(C) * 0.2}
    xq = torch.randn(B, Lq, C)
    xkv = torch.randn(B, Lkv, kv) if cross_dim else xq
    got = O._attention(sd, "a", xq, xkv, heads)
    mha = torch.nn.MultiheadAttention(C, heads, bias=True, batch_first=True, kdim=kv, vdim=kv)
    with torch.no_grad():
        if cross_dim:
            mha.q_proj_weight.copy_(sd["a.to_q.weight"])
            mha.k_proj_weight.copy_(sd["a.to_k.weight"])
            mha.v_proj_weight.copy_(sd["a.to_v.weight"])
        else:
            mha.in_proj_weight.copy_(torch.cat([sd["a.to_q.weight"], sd["a.to_k.weight"], sd["a.to_v.weight"]], 0))
        mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(sd["a.to_out.0.weight"])
        mha.out_proj.bias.copy_(sd["a.to_out.0.bias"])
        want, _ = mha(xq, xkv, xkv, need_weights=False)
    assert torch.allclose(got, want, atol=2e-6, rtol=1e-5)


def test_timestep_embedding_matches_independent_tables():
    from transformers.models.speech_to_text.modeling_speech_to_text import Speech2TextSinusoidalPositionalEmbedding
    t = torch.arange(1000)
    for dim in (64, 320):
        table = Speech2TextSinusoidalPositionalEmbedding.get_embedding(1000, dim)       # row = position = timestep
        ours = O.timestep_embedding(t, dim, False, 1.0)
        assert float((ours - table).abs().max()) < 2e-4          # sin / cos of arguments up to 999 in fp32
        # the reference's parameters: cos first, exponent / half (freq_shift 0); closed form in float64
        half = dim // 2
        freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float64) / half)
        arg = t[:, None].double() * freq[None]
        want = torch.cat([arg.cos(), arg.sin()], -1)
        ours = O.timestep_embedding(t, dim, True, 0.0)
        assert float((ours.double() - want).abs().max()) < 2e-4
        assert torch.equal(ours[:, :half], O.timestep_embedding(t, dim, False, 0.0)[:, half:])   # flip = swap halves
        assert torch.equal(ours[0], torch.cat([torch.ones(half), torch.zeros(half)]))            # t = 0: cos 1, sin 0


def test_geglu_is_value_times_erf_gelu_of_gate():
    torch.manual_seed(1)
    C = 16
    sd = {"f.net.0.proj.weight": torch.randn(8 * C, C) * 0.3, "f.net.0.proj.bias": torch.randn(8 * C) * 0.3,
          "f.net.2.weight": torch.randn(C, 4 * C) * 0.3, "f.net.2.bias": torch.randn(C) * 0.3}
    x = torch.randn(3, 5, C, dtype=torch.float32)
    got = O._feed_forward(sd, "f", x)
    xd = x.double()
    hg = xd @ sd["f.net.0.proj.weight"].double().t() + sd["f.net.0.proj.bias"].double()
    value, gate = hg[..., : 4 * C], hg[..., 4 * C:]
    act = value * (0.5 * gate * (1.0 + torch.erf(gate / math.sqrt(2.0))))          # Hendrycks & Gimpel, exact form
    want = act @ sd["f.net.2.weight"].double().t() + sd["f.net.2.bias"].double()
    assert float((got.double() - want).abs().max()) < 1e-4
    tanh_form = value * (0.5 * gate * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (gate + 0.044715 * gate ** 3))))
    assert float((act - tanh_form).abs().max()) > 1e-4             # the test can tell the two GELUs apart
    # the product packs the same halves: rows [0, 4C) values, [4C, 8C) gates (engine._pack_geglu)
    from mikudance_b200.engine import _pack_geglu
    w, b = _pack_geglu(sd["f.net.0.proj.weight"], sd["f.net.0.proj.bias"], torch.device("cpu"), 32)
    assert torch.equal(w[:16].float(), sd["f.net.0.proj.weight"][:16].half().float())                  # first 16 value rows
    assert torch.equal(w[16:32].float(), sd["f.net.0.proj.weight"][4 * C:4 * C + 16].half().float())   # their 16 gate rows
