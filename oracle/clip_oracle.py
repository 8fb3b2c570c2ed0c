"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the CLIP image-embedding stage of the pipelines
(src/pipelines/pipeline_mikudance.py:405-417): `CLIPVisionModelWithProjection(pixel_values)
.last_hidden_state` -> `vision_model.post_layernorm` applied to ALL 257 tokens -> `visual_projection`
-> image_prompt_embeds [N, 257, projection_dim].  The model itself is third-party
(`transformers`, scripts/inference_video.py:97-99); unlike diffusers it IS installed in this image
(transformers 5.5, the reference pins 4.x: same CLIP ViT architecture and state-dict keys), so this
restatement is pinned against the real implementation: tests/test_clip_oracle.py compares it with
`transformers.CLIPVisionModelWithProjection` on the same weights, and oracle/make_golden.py stores that
model's outputs as tests/golden/clip_tiny.npz for the GPU box.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

VITL14_CONFIG = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                     image_size=224, patch_size=14, projection_dim=768, layer_norm_eps=1e-5)
TINY_CLIP_CONFIG = dict(hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                        image_size=56, patch_size=14, projection_dim=64, layer_norm_eps=1e-5)


def _ln(sd, name: str, x: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"].float(), sd[name + ".bias"].float(), eps)


def _lin(sd, name: str, x: Tensor) -> Tensor:
    b = sd.get(name + ".bias")
    return F.linear(x, sd[name + ".weight"].float(), b.float() if b is not None else None)


def quick_gelu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(1.702 * x)                      # transformers activations.QuickGELUActivation


def clip_last_hidden_state(sd: Dict[str, Tensor], cfg: dict, pixel_values: Tensor) -> Tensor:
    """CLIPVisionTransformer.forward up to `last_hidden_state` (no post_layernorm): patch conv (stride =
    patch size, no bias) -> [class | patches] + position embedding -> pre_layrnorm (sic) -> encoder layers."""
    p = "vision_model."
    eps = cfg["layer_norm_eps"]
    heads = cfg["num_attention_heads"]
    x = F.conv2d(pixel_values.float(), sd[p + "embeddings.patch_embedding.weight"].float(), stride=cfg["patch_size"])
    N, C = x.shape[:2]
    x = x.flatten(2).transpose(1, 2)                                              # [N, patches, C]
    cls = sd[p + "embeddings.class_embedding"].float().expand(N, 1, C)
    x = torch.cat([cls, x], dim=1) + sd[p + "embeddings.position_embedding.weight"].float()[None]
    x = _ln(sd, p + "pre_layrnorm", x, eps)
    d = C // heads
    for i in range(cfg["num_hidden_layers"]):
        l = f"{p}encoder.layers.{i}."
        h = _ln(sd, l + "layer_norm1", x, eps)
        q = _lin(sd, l + "self_attn.q_proj", h).view(N, -1, heads, d).transpose(1, 2)
        k = _lin(sd, l + "self_attn.k_proj", h).view(N, -1, heads, d).transpose(1, 2)
        v = _lin(sd, l + "self_attn.v_proj", h).view(N, -1, heads, d).transpose(1, 2)
        a = F.scaled_dot_product_attention(q, k, v)                               # scale d^-0.5, no mask
        a = a.transpose(1, 2).reshape(N, -1, C)
        x = x + _lin(sd, l + "self_attn.out_proj", a)
        h = _ln(sd, l + "layer_norm2", x, eps)
        x = x + _lin(sd, l + "mlp.fc2", quick_gelu(_lin(sd, l + "mlp.fc1", h)))
    return x


def image_prompt_embeds(sd: Dict[str, Tensor], cfg: dict, pixel_values: Tensor) -> Tensor:
    """pipeline_mikudance.py:405-417: post_layernorm over every token, then visual_projection (no bias)."""
    x = clip_last_hidden_state(sd, cfg, pixel_values)
    x = _ln(sd, "vision_model.post_layernorm", x, cfg["layer_norm_eps"])
    return F.linear(x, sd["visual_projection.weight"].float())
