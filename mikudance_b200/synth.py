"""Synthetic (random-init) weights and inputs for benchmarks and parity tests.

There are no trained checkpoints in this environment, so bench.py / tests / smoke() use a
deterministic random-init state dict with the reference's exact key set and shapes (the contract of
`UNet3DConditionModel.state_dict()`, src/models/unet_3d_mix.py of the reference) — every tensor is
seeded by its own name, so the oracle, the reference modules and the CUDA path can all be fed
identical weights without sharing any construction order.  Motion-module `proj_out` is NOT
zero-initialised here (the reference zero-inits it at construction, src/models/motion_module.py:73-76,
which would turn the temporal path into the identity and hide it from parity tests).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Tuple

import torch

SD15_CONFIG = dict(
    in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
    block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, norm_num_groups=32, norm_eps=1e-5,
    cross_attention_dim=768, attention_head_dim=8, motion_heads=8, pe_max_len=32,
)
TINY_CONFIG = dict(SD15_CONFIG, block_out_channels=(64, 128, 256, 256), cross_attention_dim=64)


def block_plan(cfg) -> dict:
    """Static block structure (src/models/unet_3d_mix.py:124-269, unet_3d_blocks.py:654-655)."""
    boc = list(cfg["block_out_channels"])
    lpb = cfg["layers_per_block"]
    down = []
    out_c = boc[0]
    for i in range(len(boc)):
        in_c, out_c = out_c, boc[i]
        last = i == len(boc) - 1
        down.append(dict(idx=i, in_c=in_c, out_c=out_c, layers=lpb, attn=not last, downsample=not last))
    rev = list(reversed(boc))
    up = []
    out_c = rev[0]
    for i in range(len(boc)):
        prev_out = out_c
        out_c = rev[i]
        in_c = rev[min(i + 1, len(boc) - 1)]
        res_in = []
        for j in range(lpb + 1):
            res_skip = in_c if j == lpb else out_c
            res_inp = prev_out if j == 0 else out_c
            res_in.append((res_inp, res_skip))
        up.append(dict(idx=i, out_c=out_c, res_in=res_in, attn=i > 0, upsample=i < len(boc) - 1))
    return dict(down=down, mid_c=boc[-1], up=up)


def positional_encoding(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding buffer `pe` [1, max_len, d_model] (src/models/motion_module.py:275-286)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


REF_CHAR_CHANNELS = 20     # reference UNet conv_in: in_channels * 5   (src/models/unet_2d_mix.py:320-327)
REF_MOTION_CHANNELS = 2    # scene-motion map, sample[:, -2:]          (src/models/unet_2d_mix.py:1208-1209)
MAN_HIDDEN = 128           # src/models/man_module.py:14-15


def state_dict_spec(cfg, reference_unet: bool = False) -> List[Tuple[str, Tuple[int, ...], str]]:
    """[(key, shape, kind)] for every tensor of the reference UNet3DConditionModel state dict, or — with
    reference_unet=True — of the 2-D reference UNet (src/models/unet_2d_mix.py: 20-channel conv_in, one
    MANModule per down block, no motion modules, no conv_norm_out / conv_out).
    kind: 'w' weight (fan-in = prod(shape[1:])), 'b' bias, 'g' norm gain, 'pe' buffer."""
    boc = cfg["block_out_channels"]
    c0 = boc[0]
    temb = 4 * c0
    ctxd = cfg["cross_attention_dim"]
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def norm(p, c):
        spec.append((p + ".weight", (c,), "g"))
        spec.append((p + ".bias", (c,), "b"))

    def lin(p, o, i, bias=True):
        spec.append((p + ".weight", (o, i), "w"))
        if bias:
            spec.append((p + ".bias", (o,), "b"))

    def conv(p, o, i, k):
        spec.append((p + ".weight", (o, i, k, k), "w"))
        spec.append((p + ".bias", (o,), "b"))

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        lin(p + ".time_emb_proj", cout, temb)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def ff(p, c):
        lin(p + ".net.0.proj", 8 * c, c)
        lin(p + ".net.2", c, 4 * c)

    def attn(p, c, kv):
        lin(p + ".to_q", c, c, bias=False)
        lin(p + ".to_k", c, kv, bias=False)
        lin(p + ".to_v", c, kv, bias=False)
        lin(p + ".to_out.0", c, c)

    def spatial(p, c):
        norm(p + ".norm", c)
        conv(p + ".proj_in", c, c, 1)
        b = p + ".transformer_blocks.0"
        attn(b + ".attn1", c, c)
        norm(b + ".norm1", c)
        attn(b + ".attn2", c, ctxd)
        norm(b + ".norm2", c)
        ff(b + ".ff", c)
        norm(b + ".norm3", c)
        conv(p + ".proj_out", c, c, 1)

    def motion(p, c):
        t = p + ".temporal_transformer"
        norm(t + ".norm", c)
        lin(t + ".proj_in", c, c)
        b = t + ".transformer_blocks.0"
        for a in range(2):
            attn(b + f".attention_blocks.{a}", c, c)
            spec.append((b + f".attention_blocks.{a}.pos_encoder.pe", (1, cfg["pe_max_len"], c), "pe"))
            norm(b + f".norms.{a}", c)
        ff(b + ".ff", c)
        norm(b + ".ff_norm", c)
        lin(t + ".proj_out", c, c)

    if reference_unet:
        def motion(p, c):  # noqa: F811 — the 2-D reference UNet has no motion modules
            return None
    conv("conv_in", c0, REF_CHAR_CHANNELS if reference_unet else cfg["in_channels"], 3)
    lin("time_embedding.linear_1", temb, c0)
    lin("time_embedding.linear_2", temb, temb)
    plan = block_plan(cfg)
    for d in plan["down"]:
        p = f"down_blocks.{d['idx']}"
        for j in range(d["layers"]):
            resnet(f"{p}.resnets.{j}", d["in_c"] if j == 0 else d["out_c"], d["out_c"])
            if d["attn"]:
                spatial(f"{p}.attentions.{j}", d["out_c"])
            motion(f"{p}.motion_modules.{j}", d["out_c"])
        if d["downsample"]:
            conv(f"{p}.downsamplers.0.conv", d["out_c"], d["out_c"], 3)
        if reference_unet:                                  # src/models/unet_2d_mix.py:556-557
            m = f"man_blocks.{d['idx']}"
            conv(m + ".mlp_shared.0", MAN_HIDDEN, REF_MOTION_CHANNELS, 3)
            conv(m + ".mlp_gamma", d["out_c"], MAN_HIDDEN, 3)
            conv(m + ".mlp_beta", d["out_c"], MAN_HIDDEN, 3)
    mc = plan["mid_c"]
    resnet("mid_block.resnets.0", mc, mc)
    spatial("mid_block.attentions.0", mc)
    motion("mid_block.motion_modules.0", mc)
    resnet("mid_block.resnets.1", mc, mc)
    for u in plan["up"]:
        p = f"up_blocks.{u['idx']}"
        for j, (ci, cs) in enumerate(u["res_in"]):
            resnet(f"{p}.resnets.{j}", ci + cs, u["out_c"])
            if u["attn"]:
                spatial(f"{p}.attentions.{j}", u["out_c"])
            motion(f"{p}.motion_modules.{j}", u["out_c"])
        if u["upsample"]:
            conv(f"{p}.upsamplers.0.conv", u["out_c"], u["out_c"], 3)
    if not reference_unet:
        norm("conv_norm_out", c0)
        conv("conv_out", cfg["out_channels"], c0, 3)
    return spec


def _seeded_randn(name: str, shape, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return torch.randn(shape, generator=g, dtype=torch.float32)


def synthetic_state_dict(cfg, seed: int = 0, dtype=torch.float16,
                         reference_unet: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic random-init weights with the reference's key set (values rounded to `dtype`)."""
    sd = {}
    for name, shape, kind in state_dict_spec(cfg, reference_unet):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = _seeded_randn(name, shape, seed) / math.sqrt(fan_in)
        elif kind == "g":
            t = 1.0 + 0.1 * _seeded_randn(name, shape, seed)
        elif kind == "b":
            t = 0.02 * _seeded_randn(name, shape, seed)
        else:
            # the reference casts this buffer with the model (.to(fp16)); keep the key fp32 but put the
            # values on the `dtype` grid so fp32 oracle and fp16 model see the same table
            t = positional_encoding(shape[1], shape[2]).to(dtype).float()
        sd[name] = t.to(dtype) if kind != "pe" else t
    return sd


def reader_bank_order(cfg) -> List[Tuple[str, int, int]]:
    """[(attention module path, channels, downscale)] in the order ReferenceAttentionControl pairs
    reader and writer blocks (src/models/mutual_mix_attention.py:292-302,346-354)."""
    plan = block_plan(cfg)
    n = len(cfg["block_out_channels"])
    names = []
    for d in plan["down"]:
        if d["attn"]:
            for j in range(d["layers"]):
                names.append((f"down_blocks.{d['idx']}.attentions.{j}", d["out_c"], 2 ** d["idx"]))
    for u in plan["up"]:
        if u["attn"]:
            for j in range(len(u["res_in"])):
                names.append((f"up_blocks.{u['idx']}.attentions.{j}", u["out_c"], 2 ** (n - 1 - u["idx"])))
    names.append(("mid_block.attentions.0", plan["mid_c"], 2 ** (n - 1)))
    names.sort(key=lambda t: -t[1])
    return names


def synthetic_banks(cfg, n_img: int, h: int, w: int, seed: int = 102, scale: float = 1.0,
                    dtype=torch.float16) -> Dict[str, torch.Tensor]:
    """Stand-in for the reference-UNet output: one [n_img, hw, C] feature bank per spatial block
    (SURVEY.md §8d).  Rounded to fp16 like ReferenceAttentionControl.update does (:353)."""
    banks = {}
    for i, (name, c, ds) in enumerate(reader_bank_order(cfg)):
        hw = (h // ds) * (w // ds)
        banks[name] = (scale * _seeded_randn(f"bank{i}", (n_img, hw, c), seed)).to(dtype)
    return banks


def synthetic_inputs(cfg, b: int, f: int, h: int, w: int, lctx: int = 257, seed: int = 100):
    """latents [b, 4, f, h, w] and encoder_hidden_states [b, lctx, ctx_dim]; the first batch entry of
    the context is zeros when b == 2 (uncond branch, src/pipelines/pipeline_mikudance.py:418-423)."""
    sample = _seeded_randn("latents", (1, cfg["in_channels"], f, h, w), seed).repeat(b, 1, 1, 1, 1)
    ctx = _seeded_randn("ctx", (1, lctx, cfg["cross_attention_dim"]), seed + 1)
    if b == 2:
        ctx = torch.cat([torch.zeros_like(ctx), ctx], dim=0)
    return sample, ctx


def synthetic_reference_inputs(cfg, n_img: int, h: int, w: int, lctx: int = 257, seed: int = 200):
    """Inputs of the reference UNet (src/pipelines/pipeline_mikudance.py:634-653): condition latents
    [n_img, 22, h, w] (ref / skeleton / pose / face / hand latents + 2 scene-motion channels) and the
    TILED context [n_img, lctx, ctx_dim] = [uncond(zeros), cond, uncond, cond, …] (:645)."""
    x = _seeded_randn("ref_latents", (n_img, REF_CHAR_CHANNELS + REF_MOTION_CHANNELS, h, w), seed)
    c = _seeded_randn("ref_ctx", (1, lctx, cfg["cross_attention_dim"]), seed + 1)
    pair = torch.cat([torch.zeros_like(c), c], dim=0)
    ctx = pair.repeat((n_img + 1) // 2, 1, 1)[:n_img]
    return x, ctx


# ------------------------------------------------------------------------------------------------
# CLIP image encoder (transformers.CLIPVisionModelWithProjection key set) — SURVEY.md §8f row 3
# ------------------------------------------------------------------------------------------------
CLIP_VITL14_CONFIG = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                          image_size=224, patch_size=14, projection_dim=768, layer_norm_eps=1e-5)
CLIP_TINY_CONFIG = dict(hidden_size=128, intermediate_size=512, num_hidden_layers=3, num_attention_heads=2,
                        image_size=56, patch_size=14, projection_dim=64, layer_norm_eps=1e-5)


def clip_state_dict_spec(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    C, I, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["patch_size"]
    ntok = (cfg["image_size"] // P) ** 2 + 1
    v = "vision_model."
    spec: List[Tuple[str, Tuple[int, ...], str]] = [
        (v + "embeddings.class_embedding", (C,), "e"),
        (v + "embeddings.patch_embedding.weight", (C, 3, P, P), "w"),
        (v + "embeddings.position_embedding.weight", (ntok, C), "e"),
        (v + "pre_layrnorm.weight", (C,), "g"), (v + "pre_layrnorm.bias", (C,), "b"),
    ]
    for i in range(cfg["num_hidden_layers"]):
        l = f"{v}encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            spec += [(l + f"self_attn.{n}.weight", (C, C), "w"), (l + f"self_attn.{n}.bias", (C,), "b")]
        spec += [(l + "layer_norm1.weight", (C,), "g"), (l + "layer_norm1.bias", (C,), "b"),
                 (l + "mlp.fc1.weight", (I, C), "w"), (l + "mlp.fc1.bias", (I,), "b"),
                 (l + "mlp.fc2.weight", (C, I), "w"), (l + "mlp.fc2.bias", (C,), "b"),
                 (l + "layer_norm2.weight", (C,), "g"), (l + "layer_norm2.bias", (C,), "b")]
    spec += [(v + "post_layernorm.weight", (C,), "g"), (v + "post_layernorm.bias", (C,), "b"),
             ("visual_projection.weight", (cfg["projection_dim"], C), "w")]
    return spec


def synthetic_clip_state_dict(cfg, seed: int = 0, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape, kind in clip_state_dict_spec(cfg):
        if kind == "w":
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            t = _seeded_randn(name, shape, seed) / math.sqrt(fan_in)
        elif kind == "g":
            t = 1.0 + 0.1 * _seeded_randn(name, shape, seed)
        elif kind == "b":
            t = 0.02 * _seeded_randn(name, shape, seed)
        else:
            t = 0.1 * _seeded_randn(name, shape, seed)
        sd[name] = t.to(dtype)
    return sd


def synthetic_pixel_values(cfg, n: int, seed: int = 400) -> torch.Tensor:
    """CLIPImageProcessor-like input: [n, 3, image_size, image_size], roughly unit scale."""
    return _seeded_randn("clip_pixels", (n, 3, cfg["image_size"], cfg["image_size"]), seed)


# ------------------------------------------------------------------------------------------------
# VAE (diffusers.AutoencoderKL key set, SD-1.x geometry) — SURVEY.md §8f row 2
# ------------------------------------------------------------------------------------------------
SD_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                     layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215)
TINY_VAE_CONFIG = dict(SD_VAE_CONFIG, block_out_channels=(32, 64, 128, 128), norm_num_groups=8)


def vae_state_dict_spec(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    boc, lpb, lat = list(cfg["block_out_channels"]), cfg["layers_per_block"], cfg["latent_channels"]
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def norm(p, c):
        spec.extend([(p + ".weight", (c,), "g"), (p + ".bias", (c,), "b")])

    def conv(p, o, i, k):
        spec.extend([(p + ".weight", (o, i, k, k), "w"), (p + ".bias", (o,), "b")])

    def lin(p, o, i):
        spec.extend([(p + ".weight", (o, i), "w"), (p + ".bias", (o,), "b")])

    def resnet(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", co, ci, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", co, ci, 1)

    def mid(p, c):
        a = p + ".attentions.0"
        norm(a + ".group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            lin(f"{a}.{n}", c, c)
        resnet(p + ".resnets.0", c, c)
        resnet(p + ".resnets.1", c, c)

    conv("encoder.conv_in", boc[0], cfg["in_channels"], 3)
    ci = boc[0]
    for i, co in enumerate(boc):
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if i < len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        ci = co
    mid("encoder.mid_block", boc[-1])
    norm("encoder.conv_norm_out", boc[-1])
    conv("encoder.conv_out", 2 * lat, boc[-1], 3)
    conv("quant_conv", 2 * lat, 2 * lat, 1)
    conv("post_quant_conv", lat, lat, 1)
    rev = list(reversed(boc))
    conv("decoder.conv_in", rev[0], lat, 3)
    mid("decoder.mid_block", rev[0])
    ci = rev[0]
    for i, co in enumerate(rev):
        for j in range(lpb + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", ci if j == 0 else co, co)
        if i < len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        ci = co
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", cfg["out_channels"], boc[0], 3)
    return spec


def synthetic_vae_state_dict(cfg, seed: int = 0, dtype=torch.float16) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape, kind in vae_state_dict_spec(cfg):
        if kind == "w":
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            t = _seeded_randn(name, shape, seed) / math.sqrt(fan_in)
        elif kind == "g":
            t = 1.0 + 0.1 * _seeded_randn(name, shape, seed)
        else:
            t = 0.02 * _seeded_randn(name, shape, seed)
        sd[name] = t.to(dtype)
    return sd
