"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the VAE stages of the pipelines:
`vae.encode(x).latent_dist.mean` of the condition images (src/pipelines/pipeline_mikudance.py:455-549) and the
per-frame `vae.decode(z).sample` (:115-130).  The model is third-party: `diffusers.AutoencoderKL`
(`diffusers==0.24.0`, requirements.txt:7; constructed at scripts/inference_video.py:77-79) and is absent from
this image and from /root/reference, so everything below is restated from the published 0.24.0 release:
Encoder / Decoder (models/vae.py), DownEncoderBlock2D / UpDecoderBlock2D / UNetMidBlock2D
(models/unet_2d_blocks.py), ResnetBlock2D / Downsample2D / Upsample2D (models/resnet.py), the mid-block
Attention with group_norm + residual_connection (models/attention_processor.py), quant_conv / post_quant_conv
and DiagonalGaussianDistribution (models/autoencoder_kl.py).

PINNED (round 2, tests/test_vae_oracle_pin.py): diffusers itself is absent, but the installed `transformers` ships
two independent implementations of the same published latent-diffusion autoencoder — ChameleonVQVAEEncoder (LDM
Encoder) and JanusVQVAEDecoder (LDM Decoder).  With this file's seeded weights mapped key by key both halves agree
with this restatement to fp32 round-off (encoder 2e-6, decoder 1e-5 relative).  Second anchor: the parameter count
of the SD-1.x VAE it reproduces (83 653 863, tests/test_vae_oracle.py).  What stays restated-only: the names of
the diffusers state-dict keys and the 1x1 quant / post_quant convolutions (trivial).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SD_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                     layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215)
TINY_VAE_CONFIG = dict(SD_VAE_CONFIG, block_out_channels=(32, 64, 128, 128), norm_num_groups=8)

EPS = 1e-6          # resnet_eps / norm eps everywhere in the VAE


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"].float(), sd[name + ".bias"].float(), stride=stride, padding=padding)


def _gn(sd, name, x, groups):
    return F.group_norm(x, groups, sd[name + ".weight"].float(), sd[name + ".bias"].float(), EPS)


def resnet(sd, name, x, groups):
    """ResnetBlock2D with temb_channels=None: norm1-silu-conv1-norm2-silu-conv2 (+ 1x1 conv_shortcut)."""
    h = _conv(sd, name + ".conv1", F.silu(_gn(sd, name + ".norm1", x, groups)))
    h = _conv(sd, name + ".conv2", F.silu(_gn(sd, name + ".norm2", h, groups)))
    if (name + ".conv_shortcut.weight") in sd:
        x = _conv(sd, name + ".conv_shortcut", x, padding=0)
    return x + h


def mid_attention(sd, name, x, groups):
    """Attention(C, heads=1, dim_head=C, bias=True, norm_num_groups, residual_connection=True) on a 4-D input."""
    N, C, H, W = x.shape
    h = _gn(sd, name + ".group_norm", x.view(N, C, H * W), groups).transpose(1, 2)       # [N, HW, C]
    lin = lambda n, t: F.linear(t, sd[f"{name}.{n}.weight"].float(), sd[f"{name}.{n}.bias"].float())   # noqa: E731
    q, k, v = lin("to_q", h), lin("to_k", h), lin("to_v", h)
    a = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]         # one head, scale C^-0.5
    a = lin("to_out.0", a)
    return a.transpose(1, 2).reshape(N, C, H, W) + x


def mid_block(sd, name, x, groups):
    x = resnet(sd, name + ".resnets.0", x, groups)
    x = mid_attention(sd, name + ".attentions.0", x, groups)
    return resnet(sd, name + ".resnets.1", x, groups)


def encode_moments(sd: Dict[str, Tensor], cfg: dict, x: Tensor) -> Tensor:
    """AutoencoderKL.encode up to the distribution parameters: [N, 2*latent, H/8, W/8] = (mean | logvar)."""
    g, boc = cfg["norm_num_groups"], cfg["block_out_channels"]
    h = _conv(sd, "encoder.conv_in", x.float())
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"]):
            h = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g)
        if i < len(boc) - 1:                                    # Downsample2D(padding=0): pad right/bottom, stride 2
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = mid_block(sd, "encoder.mid_block", h, g)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, g)))
    return _conv(sd, "quant_conv", h, padding=0)


def encode_mean(sd, cfg, x: Tensor) -> Tensor:
    """`vae.encode(x).latent_dist.mean` (the pipelines multiply it by 0.18215 themselves)."""
    return encode_moments(sd, cfg, x)[:, :cfg["latent_channels"]]


def decode(sd: Dict[str, Tensor], cfg: dict, z: Tensor) -> Tensor:
    """`vae.decode(z).sample`: [N, latent, h, w] -> [N, 3, 8h, 8w]."""
    g, boc = cfg["norm_num_groups"], cfg["block_out_channels"]
    h = _conv(sd, "post_quant_conv", z.float(), padding=0)
    h = _conv(sd, "decoder.conv_in", h)
    h = mid_block(sd, "decoder.mid_block", h, g)
    for i in range(len(boc)):
        for j in range(cfg["layers_per_block"] + 1):
            h = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, g)
        if i < len(boc) - 1:
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, g)))
