"""RefUNetEngine — the reference UNet ("writer", src/models/unet_2d_mix.py:944-1384) as a sequence of
sm_100a kernel launches.  It is the SD-1.5 2-D UNet, so the resnet / spatial-transformer / sampler
blocks are UNetEngine's (same kernels, same NHWC token-major layout); what differs:
  * a 20-channel `conv_in` over the character-condition latents (sample[:, :-2]),
  * one MANModule after every down block, driven by the 2-channel scene-motion map (sample[:, -2:]),
  * no motion modules, no output head,
  * every transformer block's norm1 output is captured: these are the feature banks.
Its inputs are step-invariant (timestep 0, constant condition latents and embeddings), so the pipelines
run it once per context window instead of once per step per window.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops
from .engine import F16, F32, UNetEngine, _Resnet, _f32, _pack_conv3x3
from .synth import REF_CHAR_CHANNELS, REF_MOTION_CHANNELS


class RefUNetEngine(UNetEngine):
    CHAR_PAD = 24      # 20 condition channels padded to a multiple of 8 (implicit-GEMM K granularity)
    MOTION_PAD = 8

    def _pack_man(self, m):
        dev = self.dev
        o = _Resnet()
        o.c = m.mlp_gamma.out_channels
        o.w_shared = _pack_conv3x3(m.mlp_shared[0].weight, dev, cin_pad=self.MOTION_PAD)
        o.b_shared = _f32(m.mlp_shared[0].bias, dev)
        # mlp_gamma | mlp_beta share their input: one GEMM with 2C output columns
        o.w_gb = torch.cat([_pack_conv3x3(m.mlp_gamma.weight, dev), _pack_conv3x3(m.mlp_beta.weight, dev)], 0).contiguous()
        o.b_gb = torch.cat([_f32(m.mlp_gamma.bias, dev), _f32(m.mlp_beta.bias, dev)], 0).contiguous()
        return o

    def _pack(self):
        m, dev = self.model, self.dev
        temb_rows: List[torch.Tensor] = []
        temb_bias: List[torch.Tensor] = []
        assert m.conv_in.in_channels == REF_CHAR_CHANNELS
        self.conv_in_w = _pack_conv3x3(m.conv_in.weight, dev, cin_pad=self.CHAR_PAD)
        self.conv_in_b = _f32(m.conv_in.bias, dev)
        te = m.time_embedding
        f16 = lambda t: t.detach().to(device=dev, dtype=F16).contiguous()   # noqa: E731
        self.te_w1, self.te_b1 = f16(te.linear_1.weight), f16(te.linear_1.bias)
        self.te_w2, self.te_b2 = f16(te.linear_2.weight), f16(te.linear_2.bias)
        self.down, self.up = [], []
        for d, blk, man in zip(self.plan["down"], m.down_blocks, m.man_blocks):
            e = _Resnet()
            e.plan = d
            e.res = [self._pack_resnet(r, temb_rows, temb_bias) for r in blk.resnets]
            e.att = ([self._pack_spatial(a, f"down_blocks.{d['idx']}.attentions.{j}")
                      for j, a in enumerate(blk.attentions)] if d["attn"] else None)
            if d["downsample"]:
                e.ds_w = _pack_conv3x3(blk.downsamplers[0].conv.weight, dev)
                e.ds_b = _f32(blk.downsamplers[0].conv.bias, dev)
            e.man = self._pack_man(man)
            self.down.append(e)
        mid = m.mid_block
        self.mid_res = [self._pack_resnet(r, temb_rows, temb_bias) for r in mid.resnets]
        self.mid_att = self._pack_spatial(mid.attentions[0], "mid_block.attentions.0")
        for u, blk in zip(self.plan["up"], m.up_blocks):
            e = _Resnet()
            e.plan = u
            e.res = [self._pack_resnet(r, temb_rows, temb_bias) for r in blk.resnets]
            e.att = ([self._pack_spatial(a, f"up_blocks.{u['idx']}.attentions.{j}")
                      for j, a in enumerate(blk.attentions)] if u["attn"] else None)
            if u["upsample"]:
                e.us_w = _pack_conv3x3(blk.upsamplers[0].conv.weight, dev)
                e.us_b = _f32(blk.upsamplers[0].conv.bias, dev)
            self.up.append(e)
        self.temb_w = torch.cat(temb_rows, 0).contiguous()
        self.temb_b = torch.cat(temb_bias, 0).contiguous()
        edim = self.te_w1.shape[0]
        self.te_scratch = torch.empty(2 * edim + self.te_w1.shape[1], dtype=F32, device=dev)
        self.temb_vec = torch.empty(self.temb_w.shape[0], dtype=F32, device=dev)
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    # ------------------------------------------------------------------------------------------
    def _man(self, o, x, cond, N, H, W):
        """MANModule.forward (src/models/man_module.py:24-33) on NHWC rows."""
        motion = ops.cond_to_nhwc(cond, c_first=REF_CHAR_CHANNELS, c=REF_MOTION_CHANNELS, ho=H, wo=W,
                                  cpad=self.MOTION_PAD)                          # :28 nearest resize
        actv = ops.relu_(ops.gemm(motion, o.w_shared, bias=o.b_shared, conv=(N, H, W)))   # :29
        gb = ops.gemm(actv, o.w_gb, bias=o.b_gb, conv=(N, H, W))                 # :30-31
        return ops.man_modulate(x, gb, nimg=N, hw=H * W, eps=1e-5)               # :26,32

    def run(self, cond: torch.Tensor, ctx: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        """cond: [N, 22, H, W] fp16 NCHW condition latents; ctx [N or 1, L, D] fp16; the timestep is read
        from self.t_dev.  Returns (last up block's output [(N H W), C0] fp16 NHWC rows,
        banks {attention-module path: [(N hw), C] fp16})."""
        N, ctot, H, W = cond.shape
        cfg = self.cfg
        ops.time_embed(self.t_dev, self.te_w1, self.te_b1, self.te_w2, self.te_b2, self.temb_w,
                       self.temb_b, flip_sin_to_cos=bool(cfg["flip_sin_to_cos"]),
                       freq_shift=float(cfg["freq_shift"]), scratch=self.te_scratch, out=self.temb_vec)
        nctx, lctx, dctx = ctx.shape
        ctx2d = ctx.reshape(nctx * lctx, dctx)
        banks: Dict[str, torch.Tensor] = {}

        def spatial(s, x, h_, w_):
            return self._spatial(s, x, N, h_, w_, 1, ctx2d, nctx, lctx, None, 0, capture=banks)

        char = ops.cond_to_nhwc(cond, c_first=0, c=REF_CHAR_CHANNELS, ho=H, wo=W, cpad=self.CHAR_PAD)
        x = ops.gemm(char, self.conv_in_w, bias=self.conv_in_b, conv=(N, H, W))  # unet_2d_mix.py:1208-1210
        skips = [(x, H, W)]
        h_, w_ = H, W
        for e in self.down:                                                      # :1260-1289
            for j, r in enumerate(e.res):
                x = self._resnet(r, x, None, N, h_, w_)
                if e.att is not None:
                    x = spatial(e.att[j], x, h_, w_)
                skips.append((x, h_, w_))
            if e.plan["downsample"]:
                col = ops.im2col3x3(x, N, h_, w_, 2)
                h_, w_ = (h_ - 1) // 2 + 1, (w_ - 1) // 2 + 1
                x = ops.gemm(col, e.ds_w, bias=e.ds_b)
                skips.append((x, h_, w_))
            x = self._man(e.man, x, cond, N, h_, w_)     # the skips keep the un-modulated tensors (:1288-1289)
        x = self._resnet(self.mid_res[0], x, None, N, h_, w_)
        x = spatial(self.mid_att, x, h_, w_)
        x = self._resnet(self.mid_res[1], x, None, N, h_, w_)
        for e in self.up:                                                        # :1334-1368
            for j, r in enumerate(e.res):
                skip, sh, sw = skips.pop()
                assert (sh, sw) == (h_, w_), "latent size must be divisible by 8"
                x = self._resnet(r, x, skip, N, h_, w_)
                if e.att is not None:
                    x = spatial(e.att[j], x, h_, w_)
            if e.plan["upsample"]:
                x = ops.upsample2x(x, N, h_, w_)
                h_, w_ = 2 * h_, 2 * w_
                x = ops.gemm(x, e.us_w, bias=e.us_b, conv=(N, h_, w_))
        return x, banks

    # ------------------------------------------------------------------------------------------
    def forward_api(self, sample: torch.Tensor, timestep, ctx: torch.Tensor):
        """UNet2DConditionModel.forward call surface: sample [N, 22, h, w] -> (sample [N, C0, h, w],
        banks {path: [N, hw, C]})."""
        N, ctot, H, W = sample.shape
        if ctot != REF_CHAR_CHANNELS + REF_MOTION_CHANNELS:
            raise ValueError(f"the reference UNet takes {REF_CHAR_CHANNELS + REF_MOTION_CHANNELS} input "
                             f"channels (20 condition latents + 2 scene-motion), got {ctot}")
        if H % 8 or W % 8:
            raise ValueError(f"latent size {H}x{W} must be divisible by 8")
        if (H // 8) * (W // 8) <= 1:
            # nn.InstanceNorm2d raises the same way for the deepest MANModule (man_module.py:26)
            raise ValueError(f"Expected more than 1 spatial element for the deepest MANModule's instance "
                             f"norm, got latent size {H}x{W}")
        dev = self.dev
        self.set_timestep(timestep)
        s16 = sample.to(device=dev, dtype=F16).contiguous()
        c16 = ctx.to(device=dev, dtype=F16).contiguous()
        if c16.shape[0] not in (1, N):
            raise ValueError("encoder_hidden_states batch must be 1 or the sample batch")
        y, banks = self.run(s16, c16)
        c0 = y.shape[1]
        out = y.reshape(N, H, W, c0).permute(0, 3, 1, 2).contiguous().to(sample.dtype)
        return out, {k: v.reshape(N, -1, v.shape[1]) for k, v in banks.items()}
