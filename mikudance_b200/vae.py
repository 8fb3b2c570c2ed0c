"""AutoencoderKL (the SD-1.x VAE the pipelines use: `vae.encode(x).latent_dist.mean` for the condition images,
src/pipelines/pipeline_mikudance.py:455-549, and the per-frame `vae.decode(z).sample`, :115-130) on the sm_100a
kernels — SURVEY.md §8f row 2.

Same state-dict keys as `diffusers.AutoencoderKL` 0.24.0 (248 tensors, 83 653 863 parameters at SD size; the
pre-0.24 attention names query/key/value/proj_attn are mapped on load).  The modules hold parameters only; the
forward runs `VaeEngine`: every 3x3 convolution is the implicit-GEMM tcgen05 kernel (the stride-2 encoder
downsamplers go through an im2col with the right/bottom-only padding of Downsample2D(padding=0)), GroupNorm+SiLU
the fused norm kernel, the mid-block attention (ONE head of 512 channels) two GEMMs around a row-softmax kernel.
There is no PyTorch forward and no CPU path.

diffusers is absent from this image and from the reference checkout: the behaviour restated here (and in
oracle/vae_oracle.py, which the tests compare against) could not be pinned to an installed copy — see DESIGN.md.
"""
from __future__ import annotations

import json
from types import SimpleNamespace
from typing import Tuple

import torch
from torch import nn

from . import ops
from .engine import _f16, _f32, _pack_conv3x3
from .unet_3d import _Linear, _NoForward

F16 = torch.float16
F32 = torch.float32


class _Conv(nn.Conv2d):
    def forward(self, x):  # pragma: no cover
        raise RuntimeError("Conv2d: parameter container only (see AutoencoderKL.encode / decode)")


class _Resnet(_NoForward):
    def __init__(self, ci, co, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, ci, eps=1e-6)
        self.conv1 = _Conv(ci, co, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, co, eps=1e-6)
        self.conv2 = _Conv(co, co, 3, padding=1)
        self.conv_shortcut = _Conv(ci, co, 1) if ci != co else None


class _Attention(_NoForward):
    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = _Linear(c, c), _Linear(c, c), _Linear(c, c)
        self.to_out = nn.ModuleList([_Linear(c, c), nn.Dropout(0.0)])


class _Mid(_NoForward):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention(c, groups)])
        self.resnets = nn.ModuleList([_Resnet(c, c, groups), _Resnet(c, c, groups)])


class _Sampler(_NoForward):
    def __init__(self, c, stride, padding):
        super().__init__()
        self.conv = _Conv(c, c, 3, stride=stride, padding=padding)


class _Block(_NoForward):
    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()


class _Encoder(_NoForward):
    def __init__(self, cfg):
        super().__init__()
        boc, g = cfg["block_out_channels"], cfg["norm_num_groups"]
        self.conv_in = _Conv(cfg["in_channels"], boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        ci = boc[0]
        for i, co in enumerate(boc):
            b = _Block()
            for j in range(cfg["layers_per_block"]):
                b.resnets.append(_Resnet(ci if j == 0 else co, co, g))
            b.downsamplers = nn.ModuleList([_Sampler(co, 2, 0)]) if i < len(boc) - 1 else None
            self.down_blocks.append(b)
            ci = co
        self.mid_block = _Mid(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = _Conv(boc[-1], 2 * cfg["latent_channels"], 3, padding=1)


class _Decoder(_NoForward):
    def __init__(self, cfg):
        super().__init__()
        boc, g = cfg["block_out_channels"], cfg["norm_num_groups"]
        rev = list(reversed(boc))
        self.conv_in = _Conv(cfg["latent_channels"], rev[0], 3, padding=1)
        self.up_blocks = nn.ModuleList()
        self.mid_block = _Mid(rev[0], g)
        ci = rev[0]
        for i, co in enumerate(rev):
            b = _Block()
            for j in range(cfg["layers_per_block"] + 1):
                b.resnets.append(_Resnet(ci if j == 0 else co, co, g))
            b.upsamplers = nn.ModuleList([_Sampler(co, 1, 1)]) if i < len(rev) - 1 else None
            self.up_blocks.append(b)
            ci = co
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = _Conv(boc[0], cfg["out_channels"], 3, padding=1)


class DiagonalGaussianDistribution:
    """diffusers DiagonalGaussianDistribution over moments [N, 2*latent, h, w] = (mean | logvar)."""

    def __init__(self, moments: torch.Tensor):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar.float()).to(moments.dtype)
        self.var = self.std * self.std

    def sample(self, generator=None):
        gdev = generator.device if generator is not None else self.mean.device
        noise = torch.randn(self.mean.shape, generator=generator, device=gdev, dtype=self.mean.dtype)
        return self.mean + self.std * noise.to(self.mean.device)

    def mode(self):
        return self.mean


_OLD_ATTN = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * 4,
                 up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, act_fn="silu", latent_channels=4, norm_num_groups=32, sample_size=512,
                 scaling_factor=0.18215, force_upcast=True, **unused):
        super().__init__()
        boc = tuple(block_out_channels)
        ok = (tuple(down_block_types) == ("DownEncoderBlock2D",) * len(boc)
              and tuple(up_block_types) == ("UpDecoderBlock2D",) * len(boc) and act_fn in ("silu", "swish")
              and all(c % 8 == 0 and c % norm_num_groups == 0 for c in boc) and norm_num_groups <= 64
              and boc[-1] <= 1024)
        if not ok:
            raise NotImplementedError("mikudance_b200.AutoencoderKL: only the SD-1.x VAE layout is implemented")
        self._cfg = dict(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                         block_out_channels=boc, layers_per_block=layers_per_block,
                         norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self.config = SimpleNamespace(down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                                      act_fn=act_fn, sample_size=sample_size, force_upcast=force_upcast, **self._cfg)
        self.encoder = _Encoder(self._cfg)
        self.decoder = _Decoder(self._cfg)
        self.quant_conv = _Conv(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = _Conv(latent_channels, latent_channels, 1)
        self._engine = None
        self.requires_grad_(False)

    @property
    def dtype(self):
        return self.quant_conv.weight.dtype

    @property
    def device(self):
        return self.quant_conv.weight.device

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, **k):
        self._engine = None
        sd = {}
        for key, v in state_dict.items():                      # pre-0.24 attention naming, 1x1-conv shaped weights
            parts = key.split(".")
            if "attentions" in parts and parts[-2] in _OLD_ATTN:
                key = ".".join(parts[:-2] + [_OLD_ATTN[parts[-2]], parts[-1]])
                if v.dim() == 4:
                    v = v[:, :, 0, 0]
            sd[key] = v
        return super().load_state_dict(sd, strict=strict, **k)

    def engine(self):
        if self._engine is None:
            self._engine = VaeEngine(self)
        return self._engine

    @classmethod
    def from_pretrained(cls, path, subfolder=None, **unused):
        from pathlib import Path
        p = Path(path)
        if subfolder is not None:
            p = p / subfolder
        with open(p / "config.json") as fh:
            conf = {k: v for k, v in json.load(fh).items() if not k.startswith("_")}
        import inspect
        allowed = set(inspect.signature(cls.__init__).parameters) - {"self", "unused"}
        model = cls(**{k: v for k, v in conf.items() if k in allowed})
        st, bn = p / "diffusion_pytorch_model.safetensors", p / "diffusion_pytorch_model.bin"
        if st.exists():
            from safetensors.torch import load_file
            sd = load_file(str(st), device="cpu")
        elif bn.exists():
            sd = torch.load(str(bn), map_location="cpu", weights_only=True)
        else:
            raise FileNotFoundError(f"no weights file found in {p}")
        model.load_state_dict(sd)
        return model

    # -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [N, 3, H, W] -> AutoencoderKLOutput-like object with .latent_dist (mean / logvar / sample())."""
        if not x.is_cuda:
            raise RuntimeError("mikudance_b200 runs on sm_100a GPUs only (no CPU path); move the VAE and inputs to CUDA")
        dist = DiagonalGaussianDistribution(self.engine().encode_moments(x).to(x.dtype))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, **unused):
        """z [N, latent, h, w] -> DecoderOutput-like object with .sample [N, 3, 8h, 8w]."""
        if not z.is_cuda:
            raise RuntimeError("mikudance_b200 runs on sm_100a GPUs only (no CPU path); move the VAE and inputs to CUDA")
        y = self.engine().decode(z).to(z.dtype)
        return SimpleNamespace(sample=y) if return_dict else (y,)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("AutoencoderKL: use encode() / decode() (the pipelines never call forward)")


class _P:
    pass


class VaeEngine:
    """Weights packed for the kernels + the encoder / decoder as kernel launches on NHWC fp16 rows."""

    CIN_PAD = 8       # 3 image channels / 4 latent channels padded to the implicit-GEMM K granularity
    COUT_PAD = 8

    def __init__(self, model):
        p = model.quant_conv.weight
        if not p.is_cuda:
            raise RuntimeError("VaeEngine: the model must live on a CUDA (sm_100a) device; mikudance_b200 has no "
                               "CPU path")
        self._setup(model, p.device)

    # ---- packing ------------------------------------------------------------------------------
    def _conv3(self, m, cin_pad=0, cout_pad=0) -> Tuple[torch.Tensor, torch.Tensor]:
        w = _pack_conv3x3(m.weight, self.dev, cin_pad=cin_pad, cout_pad=cout_pad)
        b = torch.zeros(w.shape[0], dtype=F32, device=self.dev)
        b[: m.bias.numel()] = m.bias.detach().float()
        return w, b

    def _conv1(self, m, cin_pad=0, cout_pad=0):
        co, ci = m.weight.shape[:2]
        w = torch.zeros((max(co, cout_pad), max(ci, cin_pad)), dtype=F16, device=self.dev)
        w[:co, :ci] = m.weight.detach().to(device=self.dev, dtype=F16).reshape(co, ci)
        b = torch.zeros(w.shape[0], dtype=F32, device=self.dev)
        b[:co] = m.bias.detach().float()
        return w, b

    def _pack_resnet(self, m):
        r = _P()
        dev = self.dev
        r.n1w, r.n1b = _f16(m.norm1.weight, dev), _f16(m.norm1.bias, dev)
        r.w1, r.b1 = self._conv3(m.conv1)
        r.n2w, r.n2b = _f16(m.norm2.weight, dev), _f16(m.norm2.bias, dev)
        r.w2, r.b2 = self._conv3(m.conv2)
        r.ws = None
        if m.conv_shortcut is not None:
            r.ws, r.bs = self._conv1(m.conv_shortcut)
        return r

    def _pack_mid(self, m):
        o = _P()
        dev = self.dev
        a = m.attentions[0]
        c = a.to_q.weight.shape[0]
        o.c = c
        o.gnw, o.gnb = _f16(a.group_norm.weight, dev), _f16(a.group_norm.bias, dev)
        scale = float(c) ** -0.5                                 # one head of c channels: SDPA scale folded into q
        o.wq = (a.to_q.weight.detach().float() * scale).to(device=dev, dtype=F16).contiguous()
        o.bq = (a.to_q.bias.detach().float() * scale).to(dev).contiguous()
        o.wkv = torch.cat([_f16(a.to_k.weight, dev), _f16(a.to_v.weight, dev)], 0).contiguous()
        o.bkv = torch.cat([_f32(a.to_k.bias, dev), _f32(a.to_v.bias, dev)], 0).contiguous()
        o.wo, o.bo = _f16(a.to_out[0].weight, dev), _f32(a.to_out[0].bias, dev)
        o.res = [self._pack_resnet(r) for r in m.resnets]
        return o

    def _setup(self, model, dev):
        self.model, self.dev, self.cfg = model, dev, model._cfg
        self.groups = self.cfg["norm_num_groups"]
        lat = self.cfg["latent_channels"]
        enc, dec = model.encoder, model.decoder
        e = self.enc = _P()
        e.conv_in = self._conv3(enc.conv_in, cin_pad=self.CIN_PAD)
        e.blocks = []
        for blk in enc.down_blocks:
            b = _P()
            b.res = [self._pack_resnet(r) for r in blk.resnets]
            b.ds = self._conv3(blk.downsamplers[0].conv) if blk.downsamplers is not None else None
            e.blocks.append(b)
        e.mid = self._pack_mid(enc.mid_block)
        e.nw, e.nb = _f16(enc.conv_norm_out.weight, dev), _f16(enc.conv_norm_out.bias, dev)
        # conv_out (3x3) followed by quant_conv (1x1) is one linear map: W' = Wq Wc, b' = Wq bc + bq (exact at the
        # borders too, the 1x1 comes after the padding) -> a single 3x3 convolution with 2*latent = 8 outputs
        wq = model.quant_conv.weight.detach().float().reshape(2 * lat, 2 * lat).to(dev)
        wc = enc.conv_out.weight.detach().float().to(dev)
        fused = _Conv(wc.shape[1], 2 * lat, 3, padding=1)
        fused.weight = nn.Parameter(torch.einsum("om,mikl->oikl", wq, wc), requires_grad=False)
        fused.bias = nn.Parameter(wq @ enc.conv_out.bias.detach().float().to(dev)
                                  + model.quant_conv.bias.detach().float().to(dev), requires_grad=False)
        e.conv_out = self._conv3(fused)
        d = self.dec = _P()
        # post_quant_conv (1x1) runs as a 3x3 convolution whose only non-zero tap is the centre (it precedes
        # conv_in's zero padding, so it cannot be folded into it): same kernel path as conv_in / conv_out
        pq = _Conv(lat, lat, 3, padding=1)
        w3 = torch.zeros((lat, lat, 3, 3), dtype=F32, device=dev)
        w3[:, :, 1, 1] = model.post_quant_conv.weight.detach().float().reshape(lat, lat).to(dev)
        pq.weight = nn.Parameter(w3, requires_grad=False)
        pq.bias = nn.Parameter(model.post_quant_conv.bias.detach().float().to(dev), requires_grad=False)
        d.post_quant = self._conv3(pq, cin_pad=self.CIN_PAD, cout_pad=self.CIN_PAD)
        d.conv_in = self._conv3(dec.conv_in, cin_pad=self.CIN_PAD)
        d.mid = self._pack_mid(dec.mid_block)
        d.blocks = []
        for blk in dec.up_blocks:
            b = _P()
            b.res = [self._pack_resnet(r) for r in blk.resnets]
            b.us = self._conv3(blk.upsamplers[0].conv) if blk.upsamplers is not None else None
            d.blocks.append(b)
        d.nw, d.nb = _f16(dec.conv_norm_out.weight, dev), _f16(dec.conv_norm_out.bias, dev)
        d.conv_out = self._conv3(dec.conv_out, cout_pad=self.COUT_PAD)
        assert 2 * lat % 8 == 0, "2 * latent_channels must be a multiple of 8"

    # ---- blocks -------------------------------------------------------------------------------
    def _gn(self, x, w, b, N, hw, silu):
        return ops.groupnorm(x, w, b, nimg=N, hw=hw, groups=self.groups, eps=1e-6, silu=silu)

    def _resnet(self, r, x, N, H, W):
        h = self._gn(x, r.n1w, r.n1b, N, H * W, True)
        h = ops.gemm(h, r.w1, bias=r.b1, conv=(N, H, W))
        h = self._gn(h, r.n2w, r.n2b, N, H * W, True)
        sc = ops.gemm(x, r.ws, bias=r.bs) if r.ws is not None else x
        return ops.gemm(h, r.w2, bias=r.b2, residual=sc, conv=(N, H, W))

    def _mid(self, o, x, N, H, W):
        hw, C, dev = H * W, o.c, self.dev
        if hw % 8:
            raise NotImplementedError(f"VAE mid-block attention needs (H/8)*(W/8) % 8 == 0, got {hw} tokens")
        x = self._resnet(o.res[0], x, N, H, W)
        h = self._gn(x, o.gnw, o.gnb, N, hw, False)
        q = ops.gemm(h, o.wq, bias=o.bq)                                         # scaled queries [N*hw, C]
        k = torch.empty((N * hw, C), dtype=F16, device=dev)
        vt = torch.empty((N, C, hw), dtype=F16, device=dev)
        ops.gemm(h, o.wkv, bias=o.bkv, outs=[k, vt], trans=[False, True], trans_rows=hw)
        a = torch.empty((N * hw, C), dtype=F16, device=dev)
        for n in range(N):                                                       # one head: scores per image
            rows = slice(n * hw, (n + 1) * hw)
            s = ops.gemm(q[rows], k[rows])                                       # S = (q * scale) k^T  [hw, hw]
            ops.softmax_rows_(s)
            ops.gemm(s, vt[n], out=a[rows])                                      # O = P V   (B operand = V^T [C, hw])
        x = ops.gemm(a, o.wo, bias=o.bo, residual=x)                             # to_out + residual_connection
        return self._resnet(o.res[1], x, N, H, W)

    # ---- encoder / decoder --------------------------------------------------------------------
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        """[N, 3, H, W] -> moments [N, 2*latent, H/8, W/8] (mean | logvar), fp16."""
        N, ch, H, W = x.shape
        nd = len(self.cfg["block_out_channels"]) - 1
        if ch != self.cfg["in_channels"] or H % (1 << nd) or W % (1 << nd):
            raise ValueError(f"VAE encode: expected [N, {self.cfg['in_channels']}, H, W] with H, W divisible by "
                             f"{1 << nd}, got {tuple(x.shape)}")
        e = self.enc
        x16 = x.to(device=self.dev, dtype=F16).contiguous()
        h = ops.cond_to_nhwc(x16, c_first=0, c=ch, ho=H, wo=W, cpad=self.CIN_PAD)
        h = ops.gemm(h, e.conv_in[0], bias=e.conv_in[1], conv=(N, H, W))
        hh, ww = H, W
        for b in e.blocks:
            for r in b.res:
                h = self._resnet(r, h, N, hh, ww)
            if b.ds is not None:                                                 # Downsample2D(padding=0)
                col = ops.im2col3x3_ex(h, N, hh, ww, 2, 0)
                hh, ww = hh // 2, ww // 2
                h = ops.gemm(col, b.ds[0], bias=b.ds[1])
        h = self._mid(e.mid, h, N, hh, ww)
        h = self._gn(h, e.nw, e.nb, N, hh * ww, True)
        m = ops.gemm(h, e.conv_out[0], bias=e.conv_out[1], conv=(N, hh, ww))     # conv_out + quant_conv (folded)
        return m.view(N, hh, ww, -1).permute(0, 3, 1, 2).contiguous()

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """[N, latent, h, w] -> [N, 3, 8h, 8w], fp16."""
        N, ch, H, W = z.shape
        lat = self.cfg["latent_channels"]
        if ch != lat:
            raise ValueError(f"VAE decode: expected {lat} latent channels, got {ch}")
        d = self.dec
        z16 = z.to(device=self.dev, dtype=F16).contiguous()
        h = ops.cond_to_nhwc(z16, c_first=0, c=lat, ho=H, wo=W, cpad=self.CIN_PAD)
        h = ops.gemm(h, d.post_quant[0], bias=d.post_quant[1], conv=(N, H, W))   # post_quant_conv (centre-tap 3x3)
        h = ops.gemm(h, d.conv_in[0], bias=d.conv_in[1], conv=(N, H, W))
        h = self._mid(d.mid, h, N, H, W)
        hh, ww = H, W
        for b in d.blocks:
            for r in b.res:
                h = self._resnet(r, h, N, hh, ww)
            if b.us is not None:
                h = ops.upsample2x(h, N, hh, ww)
                hh, ww = 2 * hh, 2 * ww
                h = ops.gemm(h, b.us[0], bias=b.us[1], conv=(N, hh, ww))
        h = self._gn(h, d.nw, d.nb, N, hh * ww, True)
        y = ops.gemm(h, d.conv_out[0], bias=d.conv_out[1], conv=(N, hh, ww))
        co = self.cfg["out_channels"]
        return y[:, :co].reshape(N, hh, ww, co).permute(0, 3, 1, 2).contiguous()
