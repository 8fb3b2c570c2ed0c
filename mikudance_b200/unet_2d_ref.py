"""The reference UNet ("writer") with the reference's call surface and state-dict contract
(src/models/unet_2d_mix.py:88-1384 of Kebii/MikuDance): the SD-1.5 2-D UNet with a 20-channel
`conv_in`, one `MANModule` (src/models/man_module.py) after every down block, no output head, whose
transformer blocks record `norm1(hidden_states)` into `.bank` when a
`ReferenceAttentionControl(mode="write")` is attached (src/models/mutual_mix_attention.py:139-148).
SURVEY.md §8f row 1: the producer of the feature banks the denoising UNet reads.

Like `mikudance_b200.unet_3d`, the nn.Module tree only carries parameters under the reference's exact
names (686 SD-1.5 tensors - conv_norm_out/conv_out + 24 MAN tensors); none of the modules has a PyTorch
forward.  `forward` runs `mikudance_b200.engine_ref.RefUNetEngine` (sm_100a kernels through the C ABI);
there is no CPU / eager fallback.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Union

import torch
from torch import nn

from .synth import MAN_HIDDEN, REF_MOTION_CHANNELS, block_plan
from .unet_3d import (Attention, FeedForward, TimestepEmbedding, _Conv1x1, _LayerNorm, _Linear,
                      _NoForward)


class _Conv2d(nn.Conv2d):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("Conv2d: parameter container only (see UNet2DConditionModel.forward)")


class ResnetBlock2D(_NoForward):
    """diffusers ResnetBlock2D parameter layout (as built at src/models/unet_2d_blocks.py:547-558)."""

    def __init__(self, in_channels, out_channels, temb_channels, groups, eps):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = _Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = _Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = _Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = _Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class BasicTransformerBlock(_NoForward):
    """src/models/attention.py:12-296 (layer_norm variant).  `bank` receives norm1(hidden_states) of
    every forward while a writer ReferenceAttentionControl is attached."""

    def __init__(self, dim, heads, cross_attention_dim):
        super().__init__()
        self.norm1 = _LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads)
        self.norm2 = _LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads)
        self.norm3 = _LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.bank = []


class Transformer2DModel(_NoForward):
    """src/models/transformer_2d.py:32-250 (continuous input, conv projections)."""

    def __init__(self, heads, in_channels, cross_attention_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = _Conv1x1(in_channels, in_channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(in_channels, heads, cross_attention_dim)])
        self.proj_out = _Conv1x1(in_channels, in_channels, 1)


class MANModule(_NoForward):
    """src/models/man_module.py:8-22."""

    def __init__(self, norm_dim, m_dim, ks=3):
        super().__init__()
        self.mlp_shared = nn.Sequential(_Conv2d(m_dim, MAN_HIDDEN, ks, padding=ks // 2), nn.ReLU())
        self.mlp_gamma = _Conv2d(MAN_HIDDEN, norm_dim, ks, padding=ks // 2)
        self.mlp_beta = _Conv2d(MAN_HIDDEN, norm_dim, ks, padding=ks // 2)


class _Sampler2D(_NoForward):
    def __init__(self, channels, stride):
        super().__init__()
        self.conv = _Conv2d(channels, channels, 3, stride=stride, padding=1)


class _Block2D(_NoForward):
    """CrossAttnDownBlock2D / DownBlock2D / UNetMidBlock2DCrossAttn / UpBlock2D / CrossAttnUpBlock2D
    containers (src/models/unet_2d_blocks.py); `attentions` is registered before `resnets` there."""

    def __init__(self, has_attn):
        super().__init__()
        self.attentions = nn.ModuleList() if has_attn else None
        self.resnets = nn.ModuleList()


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class UNet2DConditionModel(nn.Module):
    """Drop-in for src.models.unet_2d_mix.UNet2DConditionModel on the inference path.  Only the SD-1.5
    configuration the reference instantiates (`from_unet(unet)` builds the class defaults,
    unet_2d_mix.py:897-920) is implemented; other options raise."""

    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                   "DownBlock2D"),
                 mid_block_type="UNetMidBlock2DCrossAttn",
                 up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 downsample_padding=1, mid_block_scale_factor=1, dropout=0.0, act_fn="silu",
                 norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=768,
                 transformer_layers_per_block=1, attention_head_dim=8, dual_cross_attention=False,
                 use_linear_projection=False, class_embed_type=None, num_class_embeds=None,
                 upcast_attention=False, resnet_time_scale_shift="default", **unused):
        super().__init__()
        cfg = dict(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                   center_input_sample=center_input_sample, flip_sin_to_cos=flip_sin_to_cos,
                   freq_shift=freq_shift, down_block_types=tuple(down_block_types),
                   mid_block_type=mid_block_type, up_block_types=tuple(up_block_types),
                   only_cross_attention=only_cross_attention, block_out_channels=tuple(block_out_channels),
                   layers_per_block=layers_per_block, downsample_padding=downsample_padding,
                   mid_block_scale_factor=mid_block_scale_factor, dropout=dropout, act_fn=act_fn,
                   norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                   cross_attention_dim=cross_attention_dim,
                   transformer_layers_per_block=transformer_layers_per_block,
                   attention_head_dim=attention_head_dim, dual_cross_attention=dual_cross_attention,
                   use_linear_projection=use_linear_projection, class_embed_type=class_embed_type,
                   num_class_embeds=num_class_embeds, upcast_attention=upcast_attention,
                   resnet_time_scale_shift=resnet_time_scale_shift)
        self.config = SimpleNamespace(**cfg)
        self._check_supported(cfg)
        self.sample_size = sample_size
        boc = tuple(block_out_channels)
        temb = boc[0] * 4
        heads = attention_head_dim
        g, eps = norm_num_groups, norm_eps
        self._plan_cfg = dict(in_channels=in_channels, out_channels=out_channels,
                              flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift,
                              block_out_channels=boc, layers_per_block=layers_per_block,
                              norm_num_groups=g, norm_eps=eps, cross_attention_dim=cross_attention_dim,
                              attention_head_dim=heads, motion_heads=8, pe_max_len=32)
        plan = block_plan(self._plan_cfg)

        self.conv_in = _Conv2d(in_channels * 5, boc[0], 3, padding=1)          # unet_2d_mix.py:320-327
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        self.man_blocks = nn.ModuleList()                                       # :480-482 registration order
        self.up_blocks = nn.ModuleList()
        self.mid_block = None
        for d in plan["down"]:
            blk = _Block2D(d["attn"])
            for j in range(d["layers"]):
                blk.resnets.append(ResnetBlock2D(d["in_c"] if j == 0 else d["out_c"], d["out_c"], temb, g, eps))
                if d["attn"]:
                    blk.attentions.append(Transformer2DModel(heads, d["out_c"], cross_attention_dim, g))
            blk.downsamplers = nn.ModuleList([_Sampler2D(d["out_c"], 2)]) if d["downsample"] else None
            self.down_blocks.append(blk)
            self.man_blocks.append(MANModule(d["out_c"], REF_MOTION_CHANNELS))  # :556-557
        mc = plan["mid_c"]
        mid = _Block2D(True)
        mid.attentions.append(Transformer2DModel(heads, mc, cross_attention_dim, g))
        mid.resnets.append(ResnetBlock2D(mc, mc, temb, g, eps))
        mid.resnets.append(ResnetBlock2D(mc, mc, temb, g, eps))
        self.mid_block = mid
        for u in plan["up"]:
            blk = _Block2D(u["attn"])
            for (ci, cs) in u["res_in"]:
                blk.resnets.append(ResnetBlock2D(ci + cs, u["out_c"], temb, g, eps))
                if u["attn"]:
                    blk.attentions.append(Transformer2DModel(heads, u["out_c"], cross_attention_dim, g))
            blk.upsamplers = nn.ModuleList([_Sampler2D(u["out_c"], 1)]) if u["upsample"] else None
            self.up_blocks.append(blk)
        self.conv_norm_out = None                                               # :677
        self._engine = None
        self._ref_control = None   # set by ReferenceAttentionControl(mode="write")
        self.requires_grad_(False)

    @staticmethod
    def _check_supported(c):
        def need(cond, what):
            if not cond:
                raise NotImplementedError(
                    f"mikudance_b200.UNet2DConditionModel (reference UNet): unsupported configuration "
                    f"({what}); only the SD-1.5 configuration of the reference's inference path is implemented")
        need(c["down_block_types"] == ("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",), "down_block_types")
        need(c["up_block_types"] == ("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3, "up_block_types")
        need(c["mid_block_type"] == "UNetMidBlock2DCrossAttn", "mid_block_type")
        need(len(c["block_out_channels"]) == 4 and c["layers_per_block"] == 2, "block_out_channels / layers")
        need(c["in_channels"] == 4, "in_channels (conv_in takes in_channels*5 = 20 condition channels)")
        need(c["act_fn"] in ("silu", "swish") and c["resnet_time_scale_shift"] == "default", "act/time norm")
        need(c["class_embed_type"] is None and c["num_class_embeds"] is None, "class embedding")
        need(not c["dual_cross_attention"] and not c["use_linear_projection"]
             and not c["only_cross_attention"] and not c["upcast_attention"], "attention variants")
        need(c["transformer_layers_per_block"] == 1 and c["dropout"] == 0.0, "transformer layers / dropout")
        need(not c["center_input_sample"] and c["mid_block_scale_factor"] == 1
             and c["downsample_padding"] == 1, "input/scale/padding")
        need(isinstance(c["attention_head_dim"], int) and isinstance(c["cross_attention_dim"], int),
             "per-block attention_head_dim / cross_attention_dim")
        for ch in c["block_out_channels"]:
            need(ch % 64 == 0 and (ch // c["attention_head_dim"]) % 8 == 0, "channel counts")

    # -------------------------------------------------------------------------------------------
    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **k):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **k)

    def engine(self):
        from .engine_ref import RefUNetEngine
        if self._engine is None:
            self._engine = RefUNetEngine(self)
        return self._engine

    def spatial_blocks(self):
        """BasicTransformerBlocks in torch_dfs order stably sorted by -width: the writer side of
        ReferenceAttentionControl's pairing (mutual_mix_attention.py:338-350)."""
        blocks = [m for m in self.modules() if isinstance(m, BasicTransformerBlock)]
        return sorted(blocks, key=lambda b: -b.norm1.normalized_shape[0])

    # -------------------------------------------------------------------------------------------
    @classmethod
    def from_unet(cls, unet):
        """src/models/unet_2d_mix.py:897-920: a new model whose conv_in carries the base UNet's 4-channel
        kernel in its first 4 input channels (zeros elsewhere) and whose time embedding / down / mid /
        up blocks are copies of the base UNet's (strict=False: the MAN blocks keep their init).
        The reference builds the class defaults (`cls(unet.config)` binds the config to `sample_size`);
        fields present on `unet.config` are honoured here, which is identical for SD-1.5."""
        import inspect
        conf = getattr(unet, "config", None) or {}
        conf = dict(conf) if isinstance(conf, dict) else dict(vars(conf))
        allowed = set(inspect.signature(cls.__init__).parameters) - {"self", "unused"}
        kw = {k: v for k, v in conf.items() if k in allowed and k not in ("down_block_types", "up_block_types",
                                                                           "mid_block_type")}
        new = cls(**kw)
        src = unet.state_dict()
        w = torch.zeros_like(new.conv_in.weight)
        w[:, :4] = src["conv_in.weight"].to(w.dtype)
        sd = {k: v for k, v in src.items()
              if k.split(".")[0] in ("time_embedding", "down_blocks", "mid_block", "up_blocks")}
        sd["conv_in.weight"] = w
        sd["conv_in.bias"] = src["conv_in.bias"]
        new.load_state_dict(sd, strict=False)
        return new

    # -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                down_intrablock_additional_residuals=None, encoder_attention_mask=None,
                return_dict: bool = True):
        """sample [N, 22, h, w] (20 condition-latent channels + 2 scene-motion channels);
        timestep scalar (the pipelines pass zeros); encoder_hidden_states [N, L, D] (one context per
        image — the pipelines pass the tiled [uncond, cond, …] tensor) or [1, L, D].
        Returns UNet2DConditionOutput(sample [N, C0, h, w]): the last up block's output (the reference
        has no output head, unet_2d_mix.py:1371-1375).  Side effect in write mode: every transformer
        block's `.bank` gets its norm1 output [N, hw, C] appended."""
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                        ("attention_mask", attention_mask), ("cross_attention_kwargs", cross_attention_kwargs),
                        ("added_cond_kwargs", added_cond_kwargs),
                        ("down_block_additional_residuals", down_block_additional_residuals),
                        ("mid_block_additional_residual", mid_block_additional_residual),
                        ("down_intrablock_additional_residuals", down_intrablock_additional_residuals),
                        ("encoder_attention_mask", encoder_attention_mask)):
            if v is not None:
                raise NotImplementedError(f"{name} is not used by the reference's pipelines and is not implemented")
        if not sample.is_cuda:
            raise RuntimeError("mikudance_b200 runs on sm_100a GPUs only (no CPU path); move the model and "
                               "inputs to CUDA")
        out, banks = self.engine().forward_api(sample, timestep, encoder_hidden_states)
        ctrl = self._ref_control
        if ctrl is not None and ctrl.get("mode") == "write":
            names = {id(m): n for n, m in self.named_modules()}
            for blk in ctrl["blocks"]:
                key = names[id(blk)].rsplit(".transformer_blocks", 1)[0]
                blk.bank.append(banks[key])                                     # mutual_mix_attention.py:140
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(sample=out)


class UNet2DWeights(nn.Module):
    """Carrier for the base SD-1.5 UNet checkpoint that scripts/inference_video.py:81-85 loads with
    `src.models.unet_2d_condition.UNet2DConditionModel.from_pretrained(path, subfolder="unet")` only to
    hand it to `UNet2DConditionModel_MIX.from_unet`: holds `config` and the state dict, has no forward."""

    def __init__(self, config: dict, state_dict: dict):
        super().__init__()
        self.config = dict(config)
        self._sd = dict(state_dict)

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def to(self, *a, **k):
        self._sd = {key: v.to(*a, **k) for key, v in self._sd.items()}
        return self

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("UNet2DWeights only carries the base checkpoint for from_unet(); the plain "
                           "SD-1.5 UNet forward is not on the MikuDance inference path")

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, **unused):
        from pathlib import Path
        path = Path(pretrained_model_path)
        if subfolder is not None:
            path = path.joinpath(subfolder)
        cfg_file = path / "config.json"
        if not cfg_file.is_file():
            raise RuntimeError(f"{cfg_file} does not exist or is not a file")
        with open(cfg_file) as fh:
            config = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
        st = path / "diffusion_pytorch_model.safetensors"
        bn = path / "diffusion_pytorch_model.bin"
        if st.exists():
            from safetensors.torch import load_file
            sd = load_file(str(st), device="cpu")
        elif bn.exists():
            sd = torch.load(str(bn), map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {path}")
        return cls(config, sd)
