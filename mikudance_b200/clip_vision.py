"""CLIP image encoder (`transformers.CLIPVisionModelWithProjection`, loaded at
scripts/inference_video.py:97-99 and used at src/pipelines/pipeline_mikudance.py:405-417) on the sm_100a
kernels — SURVEY.md §8f row 3.  One-off per clip (257 tokens of the reference image), so this is about
completeness of the native path, not about throughput.

Same state-dict keys as the transformers class (392 tensors at ViT-L/14), and the three call sites of the
pipeline work unchanged:
    emb = image_encoder(pixel_values).last_hidden_state          # encoder layers on the kernels
    emb = image_encoder.vision_model.post_layernorm(emb)          # LayerNorm kernel, every token
    image_prompt_embeds = image_encoder.visual_projection(emb)    # GEMM kernel, no bias
The modules hold parameters only; there is no PyTorch forward and no CPU path.
"""
from __future__ import annotations

import json
from types import SimpleNamespace
from typing import List

import torch
from torch import nn

from . import ops
from .unet_3d import _Linear, _NoForward

F16 = torch.float16
F32 = torch.float32


def _need_cuda16(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise RuntimeError(f"{what}: mikudance_b200 runs on sm_100a GPUs only (no CPU path)")


class KernelLayerNorm(nn.LayerNorm):
    """nn.LayerNorm parameters; forward = mdk_layernorm_f16 over the last axis."""

    def forward(self, x):
        _need_cuda16(x, "LayerNorm")
        y = ops.layernorm(x.to(F16).reshape(-1, x.shape[-1]).contiguous(), self.weight.to(F16), self.bias.to(F16),
                          eps=self.eps)
        return y.reshape(x.shape).to(x.dtype)


class KernelLinear(nn.Linear):
    """nn.Linear parameters; forward = mdk_gemm_f16."""

    def forward(self, x):
        _need_cuda16(x, "Linear")
        b = self.bias.float().contiguous() if self.bias is not None else None
        y = ops.gemm(x.to(F16).reshape(-1, x.shape[-1]).contiguous(), self.weight.to(F16).contiguous(), bias=b)
        return y.reshape(*x.shape[:-1], self.out_features).to(x.dtype)


class _PatchConv(nn.Conv2d):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("patch_embedding: parameter container only (see CLIPVisionModelWithProjection.forward)")


class CLIPVisionEmbeddings(_NoForward):
    def __init__(self, c, patch, ntok):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.zeros(c))
        self.patch_embedding = _PatchConv(3, c, patch, stride=patch, bias=False)
        self.position_embedding = nn.Embedding(ntok, c)


class CLIPAttention(_NoForward):
    def __init__(self, c):
        super().__init__()
        self.k_proj, self.v_proj = _Linear(c, c), _Linear(c, c)
        self.q_proj, self.out_proj = _Linear(c, c), _Linear(c, c)


class CLIPMLP(_NoForward):
    def __init__(self, c, inter):
        super().__init__()
        self.fc1, self.fc2 = _Linear(c, inter), _Linear(inter, c)


class CLIPEncoderLayer(_NoForward):
    def __init__(self, c, inter, eps):
        super().__init__()
        self.self_attn = CLIPAttention(c)
        self.layer_norm1 = nn.LayerNorm(c, eps=eps)
        self.mlp = CLIPMLP(c, inter)
        self.layer_norm2 = nn.LayerNorm(c, eps=eps)


class CLIPEncoder(_NoForward):
    def __init__(self, c, inter, eps, n):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(c, inter, eps) for _ in range(n)])


class CLIPVisionTransformer(_NoForward):
    def __init__(self, cfg):
        super().__init__()
        c, eps = cfg["hidden_size"], cfg["layer_norm_eps"]
        ntok = (cfg["image_size"] // cfg["patch_size"]) ** 2 + 1
        self.embeddings = CLIPVisionEmbeddings(c, cfg["patch_size"], ntok)
        self.pre_layrnorm = nn.LayerNorm(c, eps=eps)          # (sic: transformers' attribute name)
        self.encoder = CLIPEncoder(c, cfg["intermediate_size"], eps, cfg["num_hidden_layers"])
        self.post_layernorm = KernelLayerNorm(c, eps=eps)


class _Out:
    def __init__(self, last_hidden_state):
        self.last_hidden_state = last_hidden_state

    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class CLIPVisionModelWithProjection(nn.Module):
    def __init__(self, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                 image_size=224, patch_size=14, projection_dim=768, layer_norm_eps=1e-5, hidden_act="quick_gelu",
                 **unused):
        super().__init__()
        if hidden_act != "quick_gelu":
            raise NotImplementedError("only CLIP's quick_gelu MLP activation is implemented")
        if hidden_size % num_attention_heads or (hidden_size // num_attention_heads) % 8 or hidden_size % 8 \
                or hidden_size > 1536 or image_size % patch_size:
            raise NotImplementedError("unsupported CLIP vision geometry")
        self._cfg = dict(hidden_size=hidden_size, intermediate_size=intermediate_size,
                         num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                         image_size=image_size, patch_size=patch_size, projection_dim=projection_dim,
                         layer_norm_eps=layer_norm_eps)
        self.config = SimpleNamespace(hidden_act=hidden_act, **self._cfg)
        self.vision_model = CLIPVisionTransformer(self._cfg)
        self.visual_projection = KernelLinear(hidden_size, projection_dim, bias=False)
        self._engine = None
        self.requires_grad_(False)

    @property
    def dtype(self):
        return self.visual_projection.weight.dtype

    @property
    def device(self):
        return self.visual_projection.weight.device

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **k):
        self._engine = None
        sd = {key: v for key, v in state_dict.items() if not key.endswith("position_ids")}   # old checkpoints carry it
        return super().load_state_dict(sd, strict=strict, **k)

    def engine(self):
        if self._engine is None:
            self._engine = ClipEngine(self)
        return self._engine

    @classmethod
    def from_transformers(cls, model):
        """Build from a `transformers.CLIPVisionModelWithProjection` instance (config + weights)."""
        c = model.config
        new = cls(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                  num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                  image_size=c.image_size, patch_size=c.patch_size, projection_dim=c.projection_dim,
                  layer_norm_eps=c.layer_norm_eps, hidden_act=c.hidden_act)
        new.load_state_dict(model.state_dict())
        return new

    @classmethod
    def from_pretrained(cls, path, **unused):
        from pathlib import Path
        p = Path(path)
        with open(p / "config.json") as fh:
            conf = json.load(fh)
        conf = conf.get("vision_config", conf)
        import inspect
        allowed = set(inspect.signature(cls.__init__).parameters) - {"self", "unused"}
        model = cls(**{k: v for k, v in conf.items() if k in allowed})
        st, bn = p / "model.safetensors", p / "pytorch_model.bin"
        if st.exists():
            from safetensors.torch import load_file
            sd = load_file(str(st), device="cpu")
        elif bn.exists():
            sd = torch.load(str(bn), map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {p}")
        model.load_state_dict(sd)
        return model

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, **unused):
        """pixel_values [N, 3, S, S] -> object with .last_hidden_state [N, 1 + (S/P)^2, C] (before
        post_layernorm, like transformers)."""
        _need_cuda16(pixel_values, "CLIPVisionModelWithProjection")
        return _Out(self.engine().last_hidden_state(pixel_values).to(pixel_values.dtype))

    @torch.no_grad()
    def image_prompt_embeds(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """The pipelines' three calls in one: visual_projection(post_layernorm(last_hidden_state))."""
        emb = self(pixel_values).last_hidden_state
        return self.visual_projection(self.vision_model.post_layernorm(emb))


class _L:
    pass


class ClipEngine:
    """Packs the weights (fused q|k|v with bias, fp32 biases, patch kernel as a [C, 3*P*P -> pad 8] matrix,
    class token + position 0 folded into one constant row, position embedding as the patch GEMM's row bias)
    and runs the encoder as kernel launches.  Token-major fp16 activations [(n tokens), C]."""

    def __init__(self, model):
        p = model.visual_projection.weight
        if not p.is_cuda:
            raise RuntimeError("ClipEngine: the model must live on a CUDA (sm_100a) device; mikudance_b200 has "
                               "no CPU path")
        self._setup(model, p.device)

    def _setup(self, model, dev):
        self.model, self.dev, self.cfg = model, dev, model._cfg
        vm = model.vision_model
        f16 = lambda t: t.detach().to(device=dev, dtype=F16).contiguous()     # noqa: E731
        f32 = lambda t: t.detach().to(device=dev, dtype=F32).contiguous()     # noqa: E731
        C, P = self.cfg["hidden_size"], self.cfg["patch_size"]
        k = 3 * P * P
        self.kpad = (k + 7) // 8 * 8
        w = vm.embeddings.patch_embedding.weight.detach().to(device=dev, dtype=F16).reshape(C, k)
        self.patch_w = torch.zeros((C, self.kpad), dtype=F16, device=dev)
        self.patch_w[:, :k] = w                                # columns ordered (c, ky, kx) like the conv kernel
        pos = f32(vm.embeddings.position_embedding.weight)
        self.pos_patches = pos[1:].contiguous()                # fp32 row bias of the patch GEMM
        self.cls_row = (f32(vm.embeddings.class_embedding) + pos[0]).to(F16)   # class token + position 0
        self.pre_w, self.pre_b = f16(vm.pre_layrnorm.weight), f16(vm.pre_layrnorm.bias)
        self.layers: List[_L] = []
        for lyr in vm.encoder.layers:
            o = _L()
            a = lyr.self_attn
            o.wqkv = torch.cat([f16(a.q_proj.weight), f16(a.k_proj.weight), f16(a.v_proj.weight)], 0).contiguous()
            o.bqkv = torch.cat([f32(a.q_proj.bias), f32(a.k_proj.bias), f32(a.v_proj.bias)], 0).contiguous()
            o.wo, o.bo = f16(a.out_proj.weight), f32(a.out_proj.bias)
            o.ln1w, o.ln1b = f16(lyr.layer_norm1.weight), f16(lyr.layer_norm1.bias)
            o.ln2w, o.ln2b = f16(lyr.layer_norm2.weight), f16(lyr.layer_norm2.bias)
            o.w1, o.b1 = f16(lyr.mlp.fc1.weight), f32(lyr.mlp.fc1.bias)
            o.w2, o.b2 = f16(lyr.mlp.fc2.weight), f32(lyr.mlp.fc2.bias)
            self.layers.append(o)

    def last_hidden_state(self, pixel_values: torch.Tensor) -> torch.Tensor:
        cfg, dev = self.cfg, self.dev
        N, ch, H, W = pixel_values.shape
        S, P, C = cfg["image_size"], cfg["patch_size"], cfg["hidden_size"]
        if ch != 3 or H != S or W != S:
            raise ValueError(f"Input image size ({H}*{W}) doesn't match model ({S}*{S}).")
        g = S // P
        T = g * g + 1
        heads = cfg["num_attention_heads"]
        d = C // heads
        eps = cfg["layer_norm_eps"]
        # patchify: pure layout change [N, 3, g, P, g, P] -> [(N g g), (3 P P)], zero-padded to a multiple of 8
        px = pixel_values.to(device=dev, dtype=F16).reshape(N, 3, g, P, g, P).permute(0, 2, 4, 1, 3, 5)
        patches = torch.zeros((N * g * g, self.kpad), dtype=F16, device=dev)
        patches[:, : 3 * P * P] = px.reshape(N * g * g, 3 * P * P)
        emb = ops.gemm(patches, self.patch_w, row_bias=self.pos_patches, row_div=1)      # + position embedding
        x = torch.empty((N, T, C), dtype=F16, device=dev)
        x[:, 0] = self.cls_row
        x[:, 1:] = emb.view(N, g * g, C)
        x = ops.layernorm(x.view(N * T, C), self.pre_w, self.pre_b, eps=eps)
        lp = (T + 7) // 8 * 8
        for o in self.layers:
            h = ops.layernorm(x, o.ln1w, o.ln1b, eps=eps)
            q = torch.empty((N * T, C), dtype=F16, device=dev)
            k = torch.empty((N * T, C), dtype=F16, device=dev)
            vt = torch.empty((N, C, lp), dtype=F16, device=dev)
            ops.gemm(h, o.wqkv, bias=o.bqkv, outs=[q, k, vt], trans=[False, False, True], trans_rows=T)
            a = ops.attention(q, k, vt, nimg=N, lq=T, lkv=T, heads=heads, d=d)
            x = ops.gemm(a, o.wo, bias=o.bo, residual=x)
            h = ops.layernorm(x, o.ln2w, o.ln2b, eps=eps)
            f = ops.quick_gelu_(ops.gemm(h, o.w1, bias=o.b1))
            x = ops.gemm(f, o.w2, bias=o.b2, residual=x)
        return x.view(N, T, C)
