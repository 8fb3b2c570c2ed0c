"""ORACLE (test infrastructure — never imported by the product path).

CPU fp32 restatement of the reference UNet in "write" mode: the 2-D SD-1.5 UNet with a 20-channel
`conv_in` and one MANModule after every down block (src/models/unet_2d_mix.py:88-1384,
src/models/man_module.py:8-33) whose BasicTransformerBlocks record `norm1(hidden_states)` into their
`bank` (src/models/mutual_mix_attention.py:139-148).  The pipelines call it once per window per step
with timestep 0 and per-frame condition latents (src/pipelines/pipeline_mikudance.py:634-653); the
banks are what the denoising UNet's spatial self-attention adds to its keys/values.

A pure function of (state_dict, config, inputs) built from the leaf ops of oracle/unet3d_oracle.py.
Pinning: checked against the reference's own unmodified `UNet2DConditionModel` +
`ReferenceAttentionControl(mode="write")` imported from /root/reference through
oracle/diffusers_standin (tests/test_refunet_oracle.py), and against golden fixtures generated from
those modules (oracle/make_golden.py -> tests/golden/refunet_tiny_*.npz).  The diffusers pieces the
2-D blocks are built from (ResnetBlock2D, Downsample2D, Upsample2D, Attention, FeedForward,
Timesteps/TimestepEmbedding; diffusers==0.24.0, absent here) are restated from the release: for those,
parity is pinned to a restatement (oracle/README.md).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import unet3d_oracle as O

Tensor = torch.Tensor

CHAR_CHANNELS = 20      # conv_in = Conv2d(in_channels * 5, ...)   unet_2d_mix.py:320-327
MOTION_CHANNELS = 2     # sample[:, -2:] is the scene-motion map   unet_2d_mix.py:1208-1209
MAN_HIDDEN = 128        # man_module.py:14-15


def man_module(sd, name: str, x: Tensor, motion_map: Tensor) -> Tensor:
    """MANModule.forward, src/models/man_module.py:24-33: parameter-free InstanceNorm2d (eps 1e-5,
    biased variance over h*w per image and channel) modulated by gamma/beta predicted from the
    nearest-resized motion map."""
    normalized = F.instance_norm(x, eps=1e-5)                                    # :26, nn.InstanceNorm2d defaults
    m = F.interpolate(motion_map, size=x.shape[2:], mode="nearest")              # :28
    actv = O._r(F.relu(O._conv(sd, name + ".mlp_shared.0", m)))                  # :29
    gamma = O._conv(sd, name + ".mlp_gamma", actv)                               # :30
    beta = O._conv(sd, name + ".mlp_beta", actv)                                 # :31
    return O._r(normalized * (1 + gamma) + beta)                                 # :32


def transformer_2d_write(sd, name: str, x: Tensor, ctx_img: Tensor, heads: int, groups: int,
                         banks: Dict[str, Tensor]) -> Tensor:
    """Transformer2DModel.forward (src/models/transformer_2d.py:286-393, conv projections) around the
    write-mode block forward (src/models/mutual_mix_attention.py:122,139-148,243-290): the bank gets
    norm1(hidden_states); the block itself is the plain self-attn / cross-attn / GEGLU-FF block."""
    N, C, H, W = x.shape
    residual = x
    h = O._gn(sd, name + ".norm", x, groups, 1e-6)                                # transformer_2d.py norm eps 1e-6
    h = O._conv(sd, name + ".proj_in", h, padding=0)
    h = h.permute(0, 2, 3, 1).reshape(N, H * W, C)
    blk = name + ".transformer_blocks.0"
    n1 = O._ln(sd, blk + ".norm1", h)                                            # mutual_mix_attention.py:122
    banks[name] = n1.clone()                                                     # :140
    h = O._r(O._attention(sd, blk + ".attn1", n1, n1, heads) + h)                # :141-148, :245
    h = O._r(O._attention(sd, blk + ".attn2", O._ln(sd, blk + ".norm2", h), ctx_img, heads) + h)   # :247-262
    h = O._r(O._feed_forward(sd, blk + ".ff", O._ln(sd, blk + ".norm3", h)) + h)                   # :264-277
    h = h.reshape(N, H, W, C).permute(0, 3, 1, 2)
    h = O._conv(sd, name + ".proj_out", h, padding=0)
    return O._r(h + residual)


def refunet_forward(sd: Dict[str, Tensor], cfg: dict, sample: Tensor, timestep, ctx: Tensor,
                    trace: Optional[dict] = None) -> Tuple[Tensor, Dict[str, Tensor]]:
    """UNet2DConditionModel.forward of src/models/unet_2d_mix.py:944-1384 under
    ReferenceAttentionControl(mode="write", fusion_blocks="full").
    sample [N, 22, H, W] (20 character-condition channels + 2 motion channels); timestep scalar (the
    pipelines pass zeros_like(t)); ctx [N, L, D] — ONE context per image (the pipelines pass the tiled
    [uncond, cond, uncond, cond, …] tensor against a batch-major image batch, pipeline_mikudance.py:645).
    Returns (sample [N, C0, H, W] — the last up block's output: conv_norm_out / conv_out are commented
    out at :1371-1375 — and banks {attention-module path: [N, hw, C] fp32})."""
    N = sample.shape[0]
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    heads = cfg["attention_head_dim"]
    plan = O.block_plan(cfg)
    banks: Dict[str, Tensor] = {}

    t = torch.as_tensor(timestep).reshape(-1).expand(N)                           # :1070-1088
    temb = O.timestep_embedding(t, cfg["block_out_channels"][0], cfg["flip_sin_to_cos"], cfg["freq_shift"])
    emb = O._lin(sd, "time_embedding.linear_2", F.silu(O._lin(sd, "time_embedding.linear_1", temb)))
    ctx_img = ctx.float()
    if ctx_img.shape[0] == 1:
        ctx_img = ctx_img.expand(N, -1, -1)

    x = sample.float()
    char, motion = x[:, :-MOTION_CHANNELS], x[:, -MOTION_CHANNELS:]                # :1208-1209
    x = O._conv(sd, "conv_in", char)                                              # :1210
    if trace is not None:
        trace["conv_in"] = x
    skips = [x]
    for d in plan["down"]:                                                        # :1260-1289
        p = f"down_blocks.{d['idx']}"
        for j in range(d["layers"]):
            x = O.resnet_block(sd, f"{p}.resnets.{j}", x, emb, groups, eps)
            if d["attn"]:
                x = transformer_2d_write(sd, f"{p}.attentions.{j}", x, ctx_img, heads, groups, banks)
            skips.append(x)
        if d["downsample"]:
            x = O._conv(sd, f"{p}.downsamplers.0.conv", x, stride=2, padding=1)
            skips.append(x)
        # the MAN block modulates what flows on; the skip connections keep the un-modulated tensors (:1288-1289)
        x = man_module(sd, f"man_blocks.{d['idx']}", x, motion)
        if trace is not None:
            trace[f"man.{d['idx']}"] = x
    x = O.resnet_block(sd, "mid_block.resnets.0", x, emb, groups, eps)             # unet_2d_blocks.py:356-507
    x = transformer_2d_write(sd, "mid_block.attentions.0", x, ctx_img, heads, groups, banks)
    x = O.resnet_block(sd, "mid_block.resnets.1", x, emb, groups, eps)
    if trace is not None:
        trace["mid"] = x
    for u in plan["up"]:                                                          # :1334-1368
        p = f"up_blocks.{u['idx']}"
        for j in range(len(u["res_in"])):
            x = torch.cat([x, skips.pop()], dim=1)
            x = O.resnet_block(sd, f"{p}.resnets.{j}", x, emb, groups, eps)
            if u["attn"]:
                x = transformer_2d_write(sd, f"{p}.attentions.{j}", x, ctx_img, heads, groups, banks)
        if u["upsample"]:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = O._conv(sd, f"{p}.upsamplers.0.conv", x)
    return x, banks
