"""UNet3DConditionModel with the reference's call surface and state-dict contract
(src/models/unet_3d_mix.py:34-271,418-598 of Kebii/MikuDance), executed by the sm_100a kernels of
libmikudance_sm100.so through mikudance_b200.engine.UNetEngine.

The nn.Module tree below exists to carry parameters under the reference's exact names
(`down_blocks.{i}.resnets.{j}.conv1.weight`, `...attentions.{j}.transformer_blocks.0.attn1.to_q.weight`,
`...motion_modules.{j}.temporal_transformer.transformer_blocks.0.attention_blocks.{a}.pos_encoder.pe`, …
1274 tensors for SD-1.5 + motion module) so `.load_state_dict`, `.state_dict`, `.to`, `.eval` behave as
in the reference.  None of these modules has a PyTorch forward: there is no eager fallback; the
forward of the model packs the weights once (NHWC / K-major layouts, fused q|k|v and GEGLU panels)
and launches CUDA kernels only.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Union

import torch
from torch import nn

from .synth import block_plan, positional_encoding


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; the computation runs in the sm_100a "
            "kernels driven by UNet3DConditionModel.forward (no PyTorch fallback)")


class InflatedConv3d(nn.Conv2d):
    """Per-frame Conv2d weights (src/models/resnet.py:9-17); executed as implicit GEMM on tcgen05."""

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("InflatedConv3d: parameter container only (see UNet3DConditionModel.forward)")


class InflatedGroupNorm(nn.GroupNorm):
    """Per-frame GroupNorm parameters (src/models/resnet.py:20-28)."""

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("InflatedGroupNorm: parameter container only")


class _Linear(nn.Linear):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("Linear: parameter container only")


class _LayerNorm(nn.LayerNorm):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("LayerNorm: parameter container only")


class _Conv1x1(nn.Conv2d):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("Conv2d(1x1): parameter container only")


class TimestepEmbedding(_NoForward):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = _Linear(in_channels, time_embed_dim)
        self.linear_2 = _Linear(time_embed_dim, time_embed_dim)


class ResnetBlock3D(_NoForward):
    """src/models/resnet.py:123-247."""

    def __init__(self, in_channels, out_channels, temb_channels, groups, eps):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = InflatedGroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = InflatedConv3d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = _Linear(temb_channels, out_channels)
        self.norm2 = InflatedGroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = InflatedConv3d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = (InflatedConv3d(in_channels, out_channels, 1)
                              if in_channels != out_channels else None)


class Attention(_NoForward):
    """diffusers Attention parameter layout (to_q/to_k/to_v without bias, to_out.0 with bias)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8):
        super().__init__()
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = _Linear(query_dim, query_dim, bias=False)
        self.to_k = _Linear(kv, query_dim, bias=False)
        self.to_v = _Linear(kv, query_dim, bias=False)
        self.to_out = nn.ModuleList([_Linear(query_dim, query_dim), nn.Dropout(0.0)])


class _GEGLU(_NoForward):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = _Linear(dim_in, dim_out * 2)


class FeedForward(_NoForward):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim, 4 * dim), nn.Dropout(0.0), _Linear(4 * dim, dim)])


class TemporalBasicTransformerBlock(_NoForward):
    """Spatial transformer block (src/models/attention.py:298-366).  `bank` holds the reference
    features installed by ReferenceAttentionControl.update (src/models/mutual_mix_attention.py:346-354)."""

    def __init__(self, dim, heads, cross_attention_dim):
        super().__init__()
        self.attn1 = Attention(dim, None, heads)
        self.norm1 = _LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads)
        self.norm2 = _LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.norm3 = _LayerNorm(dim)
        self.bank = []


class Transformer3DModel(_NoForward):
    """src/models/transformer_3d.py:27-104."""

    def __init__(self, heads, in_channels, cross_attention_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = _Conv1x1(in_channels, in_channels, 1)
        self.transformer_blocks = nn.ModuleList(
            [TemporalBasicTransformerBlock(in_channels, heads, cross_attention_dim)])
        self.proj_out = _Conv1x1(in_channels, in_channels, 1)


class PositionalEncoding(_NoForward):
    def __init__(self, d_model, max_len):
        super().__init__()
        self.register_buffer("pe", positional_encoding(max_len, d_model))


class VersatileAttention(Attention):
    """src/models/motion_module.py:293-318."""

    def __init__(self, dim, heads, max_len):
        super().__init__(dim, None, heads)
        self.pos_encoder = PositionalEncoding(dim, max_len)


class TemporalTransformerBlock(_NoForward):
    def __init__(self, dim, heads, max_len):
        super().__init__()
        self.attention_blocks = nn.ModuleList([VersatileAttention(dim, heads, max_len) for _ in range(2)])
        self.norms = nn.ModuleList([_LayerNorm(dim) for _ in range(2)])
        self.ff = FeedForward(dim)
        self.ff_norm = _LayerNorm(dim)


class TemporalTransformer3DModel(_NoForward):
    def __init__(self, in_channels, heads, max_len, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = _Linear(in_channels, in_channels)
        self.transformer_blocks = nn.ModuleList([TemporalTransformerBlock(in_channels, heads, max_len)])
        self.proj_out = _Linear(in_channels, in_channels)


class VanillaTemporalModule(_NoForward):
    """src/models/motion_module.py:45-76 (proj_out zero-initialised at construction)."""

    def __init__(self, in_channels, heads, max_len, zero_initialize=True):
        super().__init__()
        self.temporal_transformer = TemporalTransformer3DModel(in_channels, heads, max_len)
        if zero_initialize:
            for p in self.temporal_transformer.proj_out.parameters():
                p.detach().zero_()


class _Sampler(_NoForward):
    def __init__(self, channels, stride):
        super().__init__()
        self.conv = InflatedConv3d(channels, channels, 3, stride=stride, padding=1)


class _Block(_NoForward):
    """CrossAttnDownBlock3D / DownBlock3D / CrossAttnUpBlock3D / UpBlock3D / UNetMidBlock3DCrossAttn
    containers (src/models/unet_3d_blocks.py)."""

    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = None
        self.motion_modules = nn.ModuleList()


@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


_DEFAULT_MM_KW = dict(num_attention_heads=8, num_transformer_block=1,
                      attention_block_types=("Temporal_Self", "Temporal_Self"),
                      temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                      temporal_attention_dim_div=1)


class UNet3DConditionModel(nn.Module):
    """Drop-in for src.models.unet_3d_mix.UNet3DConditionModel (same ctor kwargs, forward signature,
    state-dict keys).  Supported configuration = what the reference's inference path instantiates
    (SD-1.5 `unet/config.json` + configs/inference/mikudance_config.yaml); unsupported options raise."""

    def __init__(self, sample_size=None, in_channels=4, out_channels=4, center_input_sample=False,
                 flip_sin_to_cos=True, freq_shift=0,
                 down_block_types=("CrossAttnDownBlock3D", "CrossAttnDownBlock3D",
                                   "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type="UNetMidBlock3DCrossAttn",
                 up_block_types=("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D",
                                 "CrossAttnUpBlock3D"),
                 only_cross_attention=False, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1, act_fn="silu",
                 norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280, attention_head_dim=8,
                 dual_cross_attention=False, use_linear_projection=False, class_embed_type=None,
                 num_class_embeds=None, upcast_attention=False, resnet_time_scale_shift="default",
                 use_inflated_groupnorm=False, use_motion_module=False,
                 motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
                 motion_module_decoder_only=False, motion_module_type=None, motion_module_kwargs=None,
                 unet_use_cross_frame_attention=None, unet_use_temporal_attention=None, mode=None,
                 task_type="action", **unused):
        super().__init__()
        self._packed_device = None     # set by from_packed(): weights exist in the engine's layout only
        mmk = dict(_DEFAULT_MM_KW)
        mmk.update(dict(motion_module_kwargs or {}))
        cfg = dict(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                   center_input_sample=center_input_sample, flip_sin_to_cos=flip_sin_to_cos,
                   freq_shift=freq_shift, down_block_types=tuple(down_block_types),
                   mid_block_type=mid_block_type, up_block_types=tuple(up_block_types),
                   only_cross_attention=only_cross_attention,
                   block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                   downsample_padding=downsample_padding, mid_block_scale_factor=mid_block_scale_factor,
                   act_fn=act_fn, norm_num_groups=norm_num_groups, norm_eps=norm_eps,
                   cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                   dual_cross_attention=dual_cross_attention, use_linear_projection=use_linear_projection,
                   class_embed_type=class_embed_type, num_class_embeds=num_class_embeds,
                   upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift,
                   use_inflated_groupnorm=use_inflated_groupnorm, use_motion_module=use_motion_module,
                   motion_module_resolutions=tuple(motion_module_resolutions),
                   motion_module_mid_block=motion_module_mid_block,
                   motion_module_decoder_only=motion_module_decoder_only,
                   motion_module_type=motion_module_type, motion_module_kwargs=mmk,
                   unet_use_cross_frame_attention=unet_use_cross_frame_attention,
                   unet_use_temporal_attention=unet_use_temporal_attention, mode=mode, task_type=task_type)
        self.config = SimpleNamespace(**cfg)
        self._check_supported(cfg)
        self.sample_size = sample_size
        self.in_channels = in_channels
        self.mode = mode
        boc = tuple(block_out_channels)
        temb = boc[0] * 4
        heads = attention_head_dim if isinstance(attention_head_dim, int) else attention_head_dim[0]
        mheads = mmk["num_attention_heads"]
        max_len = mmk["temporal_position_encoding_max_len"]
        g, eps = norm_num_groups, norm_eps
        self._plan_cfg = dict(in_channels=in_channels, out_channels=out_channels,
                              flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift,
                              block_out_channels=boc, layers_per_block=layers_per_block,
                              norm_num_groups=g, norm_eps=eps, cross_attention_dim=cross_attention_dim,
                              attention_head_dim=heads, motion_heads=mheads, pe_max_len=max_len)
        plan = block_plan(self._plan_cfg)

        self.conv_in = InflatedConv3d(in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()      # registered before mid_block, as in the reference
        self.mid_block = None

        def mm(c):
            return VanillaTemporalModule(c, mheads, max_len)

        for d in plan["down"]:
            blk = _Block()
            if d["attn"]:
                blk.attentions = nn.ModuleList()
            for j in range(d["layers"]):
                blk.resnets.append(ResnetBlock3D(d["in_c"] if j == 0 else d["out_c"], d["out_c"], temb, g, eps))
                if d["attn"]:
                    blk.attentions.append(Transformer3DModel(heads, d["out_c"], cross_attention_dim, g))
                blk.motion_modules.append(mm(d["out_c"]))
            blk.downsamplers = nn.ModuleList([_Sampler(d["out_c"], 2)]) if d["downsample"] else None
            self.down_blocks.append(blk)
        mid = _Block()
        mc = plan["mid_c"]
        mid.attentions = nn.ModuleList([Transformer3DModel(heads, mc, cross_attention_dim, g)])
        mid.resnets.append(ResnetBlock3D(mc, mc, temb, g, eps))
        mid.motion_modules.append(mm(mc))
        mid.resnets.append(ResnetBlock3D(mc, mc, temb, g, eps))
        self.mid_block = mid
        for u in plan["up"]:
            blk = _Block()
            if u["attn"]:
                blk.attentions = nn.ModuleList()
            for (ci, cs) in u["res_in"]:
                blk.resnets.append(ResnetBlock3D(ci + cs, u["out_c"], temb, g, eps))
                if u["attn"]:
                    blk.attentions.append(Transformer3DModel(heads, u["out_c"], cross_attention_dim, g))
                blk.motion_modules.append(mm(u["out_c"]))
            blk.upsamplers = nn.ModuleList([_Sampler(u["out_c"], 1)]) if u["upsample"] else None
            self.up_blocks.append(blk)
        self.conv_norm_out = InflatedGroupNorm(g, boc[0], eps=eps)
        self.conv_out = InflatedConv3d(boc[0], out_channels, 3, padding=1)

        self._engine = None
        self._ref_control = None   # set by ReferenceAttentionControl(mode="read")
        self.requires_grad_(False)

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def _check_supported(c):
        def need(cond, what):
            if not cond:
                raise NotImplementedError(
                    f"mikudance_b200.UNet3DConditionModel: unsupported configuration ({what}); only the "
                    "configuration of the reference's inference path is implemented")
        need(c["down_block_types"] == ("CrossAttnDownBlock3D",) * 3 + ("DownBlock3D",), "down_block_types")
        need(c["up_block_types"] == ("UpBlock3D",) + ("CrossAttnUpBlock3D",) * 3, "up_block_types")
        need(c["mid_block_type"] == "UNetMidBlock3DCrossAttn", "mid_block_type")
        need(len(c["block_out_channels"]) == 4, "block_out_channels")
        need(c["use_inflated_groupnorm"], "use_inflated_groupnorm must be true (per-frame GroupNorm)")
        need(c["use_motion_module"] and c["motion_module_type"] == "Vanilla", "motion module")
        need(tuple(c["motion_module_resolutions"]) == (1, 2, 4, 8) and c["motion_module_mid_block"]
             and not c["motion_module_decoder_only"], "motion module placement")
        need(not c["unet_use_cross_frame_attention"] and not c["unet_use_temporal_attention"],
             "unet_use_*_attention")
        need(c["act_fn"] in ("silu", "swish") and c["resnet_time_scale_shift"] == "default", "act/time norm")
        need(c["class_embed_type"] is None and c["num_class_embeds"] is None, "class embedding")
        need(not c["dual_cross_attention"] and not c["use_linear_projection"]
             and not c["only_cross_attention"] and not c["upcast_attention"], "attention variants")
        need(not c["center_input_sample"] and c["mid_block_scale_factor"] == 1, "input/scale")
        need(isinstance(c["attention_head_dim"], int), "attention_head_dim tuple")
        mm = c["motion_module_kwargs"]
        need(mm["num_transformer_block"] == 1 and tuple(mm["attention_block_types"]) ==
             ("Temporal_Self", "Temporal_Self") and mm["temporal_position_encoding"]
             and mm["temporal_attention_dim_div"] == 1, "motion_module_kwargs")
        for ch in c["block_out_channels"]:
            need(ch % 64 == 0 and (ch // c["attention_head_dim"]) % 8 == 0, "channel counts")

    @property
    def dtype(self):
        return torch.float16 if self._packed_device is not None else self.conv_in.weight.dtype

    @property
    def device(self):
        return self._packed_device if self._packed_device is not None else self.conv_in.weight.device

    def _packed_only(self, what):
        if self._packed_device is not None:
            raise RuntimeError(f"UNet3DConditionModel.{what}: this model was built by from_packed(); its weights exist "
                               "only in the kernels' packed layout on " + str(self._packed_device) + ". Load the "
                               "original checkpoints (from_pretrained_2d + load_state_dict) to convert or move them")

    def _apply(self, fn, *a, **k):
        self._packed_only("to() / _apply")
        self._engine = None   # packed weights follow the parameters
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **k):
        self._packed_only("load_state_dict")
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **k)

    def state_dict(self, *a, **k):
        self._packed_only("state_dict")
        return super().state_dict(*a, **k)

    # -------------------------------------------------------------------------------------------
    # packed checkpoint (mikudance_b200/weight_cache.py)
    # -------------------------------------------------------------------------------------------
    def save_packed(self, path):
        """Write the engine's packed fp16 weights + this model's constructor kwargs to one safetensors file."""
        from . import weight_cache
        eng = self.engine()
        ctor = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(self.config).items()}
        weight_cache.write_file(eng, str(path), weight_cache.weights_key(self, weight_cache.layout_tag(eng)), ctor=ctor)

    @classmethod
    def from_packed(cls, path, device="cuda"):
        """A model whose module tree lives on the `meta` device (no parameter storage) and whose engine is filled
        straight from a `save_packed` file: replaces from_pretrained_2d + load_state_dict(denoising_unet.pth) +
        .to(fp16, cuda) (scripts/inference_video.py:101-114) with one read and one host->device copy."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("UNet3DConditionModel.from_packed: device must be a CUDA (sm_100a) device; "
                               "mikudance_b200 has no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        return cls._from_packed(path, dev)

    @classmethod
    def _from_packed(cls, path, dev):
        from . import weight_cache
        from .engine import UNetEngine
        with torch.device("meta"):
            model = cls(**weight_cache.read_ctor(str(path)))
        model._packed_device = dev
        eng = UNetEngine.__new__(UNetEngine)
        eng._setup(model, dev, packed_file=str(path))
        model._engine = eng
        return model

    def engine(self):
        from .engine import UNetEngine
        if self._engine is None:
            self._engine = UNetEngine(self)
        return self._engine

    def spatial_blocks(self):
        """Transformer blocks in torch_dfs order stably sorted by -width: the order
        ReferenceAttentionControl uses to pair reader and writer (mutual_mix_attention.py:292-302)."""
        blocks = []
        for m in self.modules():          # pre-order DFS == reference torch_dfs
            if isinstance(m, TemporalBasicTransformerBlock):
                blocks.append(m)
        return sorted(blocks, key=lambda b: -b.norm1.normalized_shape[0])

    # -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, class_labels=None, attention_mask=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                return_dict: bool = True, self_attention_additional_feats=None):
        """sample [B, 4, f, h, w]; timestep scalar (or a tensor whose entries are all equal);
        encoder_hidden_states [B, L, D].  Returns UNet3DConditionOutput(sample [B, 4, f, h, w])."""
        if class_labels is not None or attention_mask is not None or \
                down_block_additional_residuals is not None or mid_block_additional_residual is not None:
            raise NotImplementedError("class_labels / attention_mask / additional residuals are not "
                                      "used by the reference's denoising loop and are not implemented")
        if not sample.is_cuda:
            raise RuntimeError("mikudance_b200 runs on sm_100a GPUs only (no CPU path); move the model and "
                               "inputs to CUDA")
        eng = self.engine()
        out = eng.forward_api(sample, timestep, encoder_hidden_states)
        if not return_dict:
            return (out,)
        return UNet3DConditionOutput(sample=out)

    # -------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, motion_module_path, subfolder=None,
                           unet_additional_kwargs=None, mm_zero_proj_out=False):
        """src/models/unet_3d_mix.py:600-691: read SD-1.5 `config.json`, build, load the 2-D weights
        (.safetensors or .bin) then the motion-module checkpoint, strict=False."""
        from pathlib import Path
        path = Path(pretrained_model_path)
        if subfolder is not None:
            path = path.joinpath(subfolder)
        cfg_file = path / "config.json"
        if not cfg_file.is_file():
            raise RuntimeError(f"{cfg_file} does not exist or is not a file")
        with open(cfg_file) as fh:
            unet_config = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
        unet_config["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
        unet_config["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
        unet_config["mid_block_type"] = "UNetMidBlock3DCrossAttn"
        import inspect
        allowed = set(inspect.signature(cls.__init__).parameters)
        kwargs = {k: v for k, v in unet_config.items() if k in allowed}
        kwargs.update(dict(unet_additional_kwargs or {}))
        model = cls(**kwargs)
        st = path / "diffusion_pytorch_model.safetensors"
        bn = path / "diffusion_pytorch_model.bin"
        if st.exists():
            from safetensors.torch import load_file
            state_dict = load_file(str(st), device="cpu")
        elif bn.exists():
            state_dict = torch.load(str(bn), map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {path}")
        mpath = Path(motion_module_path)
        if mpath.exists() and mpath.is_file():
            if mpath.suffix.lower() in (".pth", ".pt", ".ckpt"):
                motion_sd = torch.load(str(mpath), map_location="cpu", weights_only=True)
            elif mpath.suffix.lower() == ".safetensors":
                from safetensors.torch import load_file
                motion_sd = load_file(str(mpath), device="cpu")
            else:
                raise RuntimeError(f"unknown file format for motion module weights: {mpath.suffix}")
            if mm_zero_proj_out:
                motion_sd = {k: v for k, v in motion_sd.items() if "proj_out" not in k}
            state_dict.update(motion_sd)
        model.load_state_dict(state_dict, strict=False)
        return model
