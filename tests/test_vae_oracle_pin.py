"""CPU: PIN of oracle/vae_oracle.py to third-party code that is installed here.

diffusers (whose AutoencoderKL the reference instantiates, scripts/inference_video.py:77-79) is absent from this
image, but `transformers` ships two independent implementations of the SAME published architecture — the
latent-diffusion / taming-transformers autoencoder that SD's VAE and diffusers' AutoencoderKL re-express:
  * transformers.models.chameleon.modeling_chameleon.ChameleonVQVAEEncoder   (LDM `Encoder`: conv_in, per level
    num_res_blocks resnets + stride-2 conv with (0,1,0,1) padding, mid resnet-attention-resnet, GroupNorm(32,
    eps 1e-6) + swish + conv_out with 2 x z channels)
  * transformers.models.janus.modeling_janus.JanusVQVAEDecoder               (LDM `Decoder`: conv_in, mid block,
    per level num_res_blocks + 1 resnets + nearest x2 + conv, GroupNorm + swish + conv_out; its extra attention
    blocks in the lowest-resolution level have no counterpart in SD's decoder and are turned into exact identities
    by zeroing their proj_out — x + proj_out(...) = x)
The oracle's seeded weights are mapped key by key (diffusers names -> LDM names: resnets.j -> block[j],
downsamplers.0.conv -> downsample.conv, conv_shortcut -> nin_shortcut, attention Linear [C, C] -> 1x1 conv
[C, C, 1, 1], conv_norm_out -> norm_out) and both halves must agree with the oracle to fp32 round-off.  The 1x1
quant_conv / post_quant_conv of AutoencoderKL (absent from a VQ model) are applied with torch.nn.functional.conv2d.
Both third-party models hard-code 32 GroupNorm groups, hence the 32-multiple channel widths of the pin config."""
import pytest
import torch
import torch.nn.functional as F

from mikudance_b200 import synth
from oracle import vae_oracle as V

PIN_CFG = dict(V.SD_VAE_CONFIG, block_out_channels=(32, 64, 128, 128), norm_num_groups=32)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _map_resnet(dst, src_sd, src):
    out = {}
    for n in ("norm1", "conv1", "norm2", "conv2"):
        for s in ("weight", "bias"):
            out[f"{dst}.{n}.{s}"] = src_sd[f"{src}.{n}.{s}"]
    if f"{src}.conv_shortcut.weight" in src_sd:
        out[f"{dst}.nin_shortcut.weight"] = src_sd[f"{src}.conv_shortcut.weight"]
        out[f"{dst}.nin_shortcut.bias"] = src_sd[f"{src}.conv_shortcut.bias"]
    return out


def _map_attn(dst, src_sd, src):
    out = {f"{dst}.norm.weight": src_sd[f"{src}.group_norm.weight"], f"{dst}.norm.bias": src_sd[f"{src}.group_norm.bias"]}
    for a, b in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
        out[f"{dst}.{a}.weight"] = src_sd[f"{src}.{b}.weight"][:, :, None, None]
        out[f"{dst}.{a}.bias"] = src_sd[f"{src}.{b}.bias"]
    return out


@pytest.fixture(scope="module")
def weights():
    return {k: v.float() for k, v in synth.synthetic_vae_state_dict(PIN_CFG, seed=5).items()}


def test_encoder_equals_transformers_ldm_encoder(weights):
    cham = pytest.importorskip("transformers.models.chameleon.modeling_chameleon")
    from transformers.models.chameleon.configuration_chameleon import ChameleonVQVAEConfig
    boc = PIN_CFG["block_out_channels"]
    cfg = ChameleonVQVAEConfig(double_latent=True, latent_channels=4, in_channels=3, base_channels=boc[0],
                               channel_multiplier=[c // boc[0] for c in boc], num_res_blocks=2, attn_resolutions=None,
                               dropout=0.0, attn_type="vanilla", resolution=64)
    enc = cham.ChameleonVQVAEEncoder(cfg).eval()
    sd = {}
    for s in ("weight", "bias"):
        sd[f"conv_in.{s}"] = weights[f"encoder.conv_in.{s}"]
        sd[f"norm_out.{s}"] = weights[f"encoder.conv_norm_out.{s}"]
        sd[f"conv_out.{s}"] = weights[f"encoder.conv_out.{s}"]
    for i in range(len(boc)):
        for j in range(2):
            sd.update(_map_resnet(f"down.{i}.block.{j}", weights, f"encoder.down_blocks.{i}.resnets.{j}"))
        if i < len(boc) - 1:
            for s in ("weight", "bias"):
                sd[f"down.{i}.downsample.conv.{s}"] = weights[f"encoder.down_blocks.{i}.downsamplers.0.conv.{s}"]
    sd.update(_map_resnet("mid.block_1", weights, "encoder.mid_block.resnets.0"))
    sd.update(_map_resnet("mid.block_2", weights, "encoder.mid_block.resnets.1"))
    sd.update(_map_attn("mid.attn_1", weights, "encoder.mid_block.attentions.0"))
    res = enc.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    x = synth._seeded_randn("vae_pin_img", (2, 3, 64, 96), 3)
    with torch.no_grad():
        third = F.conv2d(enc(x.clone()), weights["quant_conv.weight"], weights["quant_conv.bias"])
        ours = V.encode_moments(weights, PIN_CFG, x)
    assert ours.shape == third.shape == (2, 8, 8, 12)
    assert _rel(ours, third) < 1e-5, _rel(ours, third)      # fp32 round-off (x*sigmoid vs silu, bmm vs SDPA)
    assert torch.equal(V.encode_mean(weights, PIN_CFG, x), ours[:, :4])


def test_decoder_equals_transformers_ldm_decoder(weights):
    janus = pytest.importorskip("transformers.models.janus.modeling_janus")
    from transformers.models.janus.configuration_janus import JanusVQVAEConfig
    boc = PIN_CFG["block_out_channels"]
    cfg = JanusVQVAEConfig(double_latent=False, latent_channels=4, in_channels=3, out_channels=3, base_channels=boc[0],
                           channel_multiplier=[c // boc[0] for c in boc], num_res_blocks=2, dropout=0.0)
    dec = janus.JanusVQVAEDecoder(cfg).eval()
    n = len(boc)
    sd = {}
    for s in ("weight", "bias"):
        sd[f"conv_in.{s}"] = weights[f"decoder.conv_in.{s}"]
        sd[f"norm_out.{s}"] = weights[f"decoder.conv_norm_out.{s}"]
        sd[f"conv_out.{s}"] = weights[f"decoder.conv_out.{s}"]
    sd.update(_map_resnet("mid.block_1", weights, "decoder.mid_block.resnets.0"))
    sd.update(_map_resnet("mid.block_2", weights, "decoder.mid_block.resnets.1"))
    sd.update(_map_attn("mid.attn_1", weights, "decoder.mid_block.attentions.0"))
    for i in range(n):                                   # Janus' up[i] runs in the order of diffusers' up_blocks[i]
        for j in range(3):
            sd.update(_map_resnet(f"up.{i}.block.{j}", weights, f"decoder.up_blocks.{i}.resnets.{j}"))
        if i < n - 1:
            for s in ("weight", "bias"):
                sd[f"up.{i}.upsample.conv.{s}"] = weights[f"decoder.up_blocks.{i}.upsamplers.0.conv.{s}"]
    own = dec.state_dict()
    for k, v in own.items():                             # the attention blocks SD's decoder does not have: identity
        if ".attn." in k and k.startswith("up."):
            sd[k] = torch.zeros_like(v) if ".proj_out." in k else v
    res = dec.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    z = 0.5 * synth._seeded_randn("vae_pin_lat", (2, 4, 8, 12), 4)
    with torch.no_grad():
        third = dec(F.conv2d(z, weights["post_quant_conv.weight"], weights["post_quant_conv.bias"]))
        ours = V.decode(weights, PIN_CFG, z)
    assert ours.shape == third.shape == (2, 3, 64, 96)
    assert _rel(ours, third) < 1e-5, _rel(ours, third)      # fp32 round-off (x*sigmoid vs silu, bmm vs SDPA)
