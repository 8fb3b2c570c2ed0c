"""UNetEngine — packs a UNet3DConditionModel's weights into kernel layouts and runs the forward as a
sequence of sm_100a kernel launches (mikudance_b200.ops -> C ABI).  One canonical activation layout
end to end: fp16 tokens [(b f) * h * w, C] ("NHWC"); the reference's `b c f h w <-> (b f) c h w`
permute copies around every conv / norm (src/models/resnet.py:13-15,24-26) do not exist here.

Frame sharding: with a process group, each rank runs the UNet on its slice of the
window's frames (one CFG branch per half of the ranks, or both on every rank); the only exchange is inside each
motion module: a frames<->pixels all-to-all before and after the temporal transformer (default; `_motion_a2a`,
no layout copies) or, alternatively, an NCCL all-gather of the temporal attention's K/V (`MDK_SHARD_MODE=allgather`).
torch.distributed is the plumbing; both are capturable in CUDA graphs.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .synth import block_plan

F16 = torch.float16
F32 = torch.float32


def _f16(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=F16).contiguous()


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=F32).contiguous()


def _pack_conv3x3(w: torch.Tensor, dev, cin_pad: int = 0, cout_pad: int = 0) -> torch.Tensor:
    """OIHW [Cout, Cin, 3, 3] -> [Cout, (kh, kw, Cin)] so the implicit-GEMM K axis is contiguous."""
    w = w.detach().to(device=dev, dtype=F16)
    co, ci = w.shape[:2]
    if cin_pad > ci:
        w = torch.cat([w, w.new_zeros(co, cin_pad - ci, 3, 3)], 1)
    if cout_pad > co:
        w = torch.cat([w, w.new_zeros(cout_pad - co, w.shape[1], 3, 3)], 0)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _pack_geglu(w: torch.Tensor, b: torch.Tensor, dev, block: int):
    """Linear(C, 8C) of GEGLU: rows [0,4C) values, [4C,8C) gates -> panels of `block` rows holding
    [block/2 value rows | the matching block/2 gate rows] (epilogue computes value*gelu(gate))."""
    w = w.detach().to(device=dev, dtype=F16)
    b = b.detach().to(device=dev, dtype=F32)
    n2 = w.shape[0] // 2
    half = block // 2
    assert n2 % half == 0, (w.shape, block)
    idx = []
    for t in range(n2 // half):
        idx.append(torch.arange(t * half, (t + 1) * half))
        idx.append(torch.arange(n2 + t * half, n2 + (t + 1) * half))
    idx = torch.cat(idx).to(dev)
    return w[idx].contiguous(), b[idx].contiguous()


class _Resnet:
    pass


class UNetEngine:
    def __init__(self, model):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("UNetEngine: the model must live on a CUDA (sm_100a) device; "
                               "mikudance_b200 has no CPU path")
        if p.dtype != F16:
            raise RuntimeError(f"UNetEngine: weights must be float16 (got {p.dtype}); call "
                               ".to(dtype=torch.float16) like scripts/inference_video.py does")
        self._setup(model, p.device)

    def _setup(self, model, dev, packed_file=None):
        self.model = model
        self.dev = dev
        self.cfg = model._plan_cfg
        self.plan = block_plan(self.cfg)
        self.groups = self.cfg["norm_num_groups"]
        self.eps = self.cfg["norm_eps"]
        self.heads = self.cfg["attention_head_dim"]
        self.mheads = self.cfg["motion_heads"]
        from ._lib import load_library
        self.geglu_block = int(load_library().mdk_gemm_geglu_block())
        self.trace = None       # optional {stage: (tensor, N, h, w)} of intermediate activations (tests)
        self.pg = None          # torch.distributed process group for frame sharding
        import os
        # "a2a": two all-to-alls per motion module (default); "allgather": K/V all-gather per attention
        self.shard_mode = os.environ.get("MDK_SHARD_MODE", "a2a")
        self.force_simt = False   # tests: run the gathered-layout temporal kernel on one GPU
        self.world = 1
        self.rank = 0
        if packed_file is not None:      # UNet3DConditionModel.from_packed: the module holds no weights
            from . import weight_cache
            if not weight_cache.read_file(self, packed_file):
                raise RuntimeError(f"{packed_file}: packed for another kernel layout / GEGLU panel width; "
                                   "re-create it with save_packed() from the original checkpoints")
            return
        from .weight_cache import PackedWeightCache
        cache = PackedWeightCache.from_env()        # MDK_WEIGHT_CACHE=<dir>
        key = cache.load(self) if cache is not None else None
        self.weight_cache = cache.last if cache is not None else None
        if self.weight_cache != "hit":
            before = set(vars(self))
            self._pack()
            self._packed_attrs = sorted(set(vars(self)) - before)
            if cache is not None:
                cache.store(self, key)

    # ------------------------------------------------------------------------------------------
    # weight packing
    # ------------------------------------------------------------------------------------------
    def _pack_resnet(self, m, temb_rows: List[torch.Tensor], temb_bias: List[torch.Tensor]):
        dev = self.dev
        r = _Resnet()
        r.cin, r.cout = m.in_channels, m.out_channels
        r.n1w, r.n1b = _f16(m.norm1.weight, dev), _f16(m.norm1.bias, dev)
        r.w1 = _pack_conv3x3(m.conv1.weight, dev)
        r.temb_off = sum(t.shape[0] for t in temb_rows)
        temb_rows.append(_f16(m.time_emb_proj.weight, dev))
        temb_bias.append(_f32(m.time_emb_proj.bias, dev) + _f32(m.conv1.bias, dev))
        r.n2w, r.n2b = _f16(m.norm2.weight, dev), _f16(m.norm2.bias, dev)
        r.w2 = _pack_conv3x3(m.conv2.weight, dev)
        r.b2 = _f32(m.conv2.bias, dev)
        if m.conv_shortcut is not None:
            r.ws = _f16(m.conv_shortcut.weight.reshape(r.cout, r.cin), dev)
            r.bs = _f32(m.conv_shortcut.bias, dev)
        else:
            r.ws = None
        return r

    def _pack_ff(self, ff, o):
        o.ff1_w, o.ff1_b = _pack_geglu(ff.net[0].proj.weight, ff.net[0].proj.bias, self.dev,
                                       self.geglu_block)
        o.ff2_w, o.ff2_b = _f16(ff.net[2].weight, self.dev), _f32(ff.net[2].bias, self.dev)

    def _pack_spatial(self, m, name):
        dev = self.dev
        s = _Resnet()
        s.name = name
        s.block = m.transformer_blocks[0]
        s.c = m.norm.num_channels
        s.gnw, s.gnb = _f16(m.norm.weight, dev), _f16(m.norm.bias, dev)
        s.pin_w = _f16(m.proj_in.weight.reshape(s.c, s.c), dev)
        s.pin_b = _f32(m.proj_in.bias, dev)
        s.pout_w = _f16(m.proj_out.weight.reshape(s.c, s.c), dev)
        s.pout_b = _f32(m.proj_out.bias, dev)
        b = m.transformer_blocks[0]
        for i, ln in ((1, b.norm1), (2, b.norm2), (3, b.norm3)):
            setattr(s, f"ln{i}w", _f16(ln.weight, dev))
            setattr(s, f"ln{i}b", _f16(ln.bias, dev))
        a1 = b.attn1
        s.wq = _f16(a1.to_q.weight, dev)
        s.wkv = torch.cat([_f16(a1.to_k.weight, dev), _f16(a1.to_v.weight, dev)], 0).contiguous()
        s.wqkv = torch.cat([s.wq, s.wkv], 0).contiguous()
        s.wo, s.bo = _f16(a1.to_out[0].weight, dev), _f32(a1.to_out[0].bias, dev)
        a2 = b.attn2
        s.wq2 = _f16(a2.to_q.weight, dev)
        s.wkv2 = torch.cat([_f16(a2.to_k.weight, dev), _f16(a2.to_v.weight, dev)], 0).contiguous()
        s.wo2, s.bo2 = _f16(a2.to_out[0].weight, dev), _f32(a2.to_out[0].bias, dev)
        self._pack_ff(b.ff, s)
        return s

    def _pack_motion(self, m):
        dev = self.dev
        t = m.temporal_transformer
        o = _Resnet()
        o.c = t.norm.num_channels
        o.gnw, o.gnb = _f16(t.norm.weight, dev), _f16(t.norm.bias, dev)
        o.pin_w, o.pin_b = _f16(t.proj_in.weight, dev), _f32(t.proj_in.bias, dev)
        o.pout_w, o.pout_b = _f16(t.proj_out.weight, dev), _f32(t.proj_out.bias, dev)
        blk = t.transformer_blocks[0]
        o.att = []
        for a in range(2):
            ab = blk.attention_blocks[a]
            e = _Resnet()
            e.lnw, e.lnb = _f16(blk.norms[a].weight, dev), _f16(blk.norms[a].bias, dev)
            wq = _f16(ab.to_q.weight, dev)
            e.wqkv = torch.cat([wq, _f16(ab.to_k.weight, dev), _f16(ab.to_v.weight, dev)], 0).contiguous()
            e.wo, e.bo = _f16(ab.to_out[0].weight, dev), _f32(ab.to_out[0].bias, dev)
            # PE enters the query only (motion_module.py:404-417): (x + pe) Wq^T = x Wq^T + pe Wq^T.
            # The table pe @ Wq^T is a constant of the weights: computed once here in fp32 on the GPU.
            # It rides on the fused q|k|v GEMM as a per-frame row bias (zero for the k and v columns), so
            # the query gets it in fp32 before its single fp16 rounding.
            pe = ab.pos_encoder.pe.detach().to(device=dev, dtype=F32)[0]
            e.pe_qkv = torch.zeros((pe.shape[0], 3 * o.c), dtype=F32, device=dev)
            e.pe_qkv[:, : o.c] = pe @ wq.float().t()
            o.att.append(e)
        o.ffnw, o.ffnb = _f16(blk.ff_norm.weight, dev), _f16(blk.ff_norm.bias, dev)
        self._pack_ff(blk.ff, o)
        return o

    def _pack(self):
        m, dev = self.model, self.dev
        temb_rows: List[torch.Tensor] = []
        temb_bias: List[torch.Tensor] = []
        self.cin_pad = 8
        self.cout_pad = 8
        self.conv_in_w = _pack_conv3x3(m.conv_in.weight, dev, cin_pad=self.cin_pad)
        self.conv_in_b = _f32(m.conv_in.bias, dev)
        te = m.time_embedding
        self.te_w1, self.te_b1 = _f16(te.linear_1.weight, dev), _f16(te.linear_1.bias, dev)
        self.te_w2, self.te_b2 = _f16(te.linear_2.weight, dev), _f16(te.linear_2.bias, dev)
        self.down, self.up = [], []
        for d, blk in zip(self.plan["down"], m.down_blocks):
            e = _Resnet()
            e.plan = d
            e.res = [self._pack_resnet(r, temb_rows, temb_bias) for r in blk.resnets]
            e.att = ([self._pack_spatial(a, f"down_blocks.{d['idx']}.attentions.{j}")
                      for j, a in enumerate(blk.attentions)] if d["attn"] else None)
            e.mm = [self._pack_motion(x) for x in blk.motion_modules]
            if d["downsample"]:
                e.ds_w = _pack_conv3x3(blk.downsamplers[0].conv.weight, dev)
                e.ds_b = _f32(blk.downsamplers[0].conv.bias, dev)
            self.down.append(e)
        mid = m.mid_block
        self.mid_res = [self._pack_resnet(r, temb_rows, temb_bias) for r in mid.resnets]
        self.mid_att = self._pack_spatial(mid.attentions[0], "mid_block.attentions.0")
        self.mid_mm = self._pack_motion(mid.motion_modules[0])
        for u, blk in zip(self.plan["up"], m.up_blocks):
            e = _Resnet()
            e.plan = u
            e.res = [self._pack_resnet(r, temb_rows, temb_bias) for r in blk.resnets]
            e.att = ([self._pack_spatial(a, f"up_blocks.{u['idx']}.attentions.{j}")
                      for j, a in enumerate(blk.attentions)] if u["attn"] else None)
            e.mm = [self._pack_motion(x) for x in blk.motion_modules]
            if u["upsample"]:
                e.us_w = _pack_conv3x3(blk.upsamplers[0].conv.weight, dev)
                e.us_b = _f32(blk.upsamplers[0].conv.bias, dev)
            self.up.append(e)
        self.out_nw, self.out_nb = _f16(m.conv_norm_out.weight, dev), _f16(m.conv_norm_out.bias, dev)
        self.conv_out_w = _pack_conv3x3(m.conv_out.weight, dev, cout_pad=self.cout_pad)
        ob = torch.zeros(self.cout_pad, dtype=F32, device=dev)
        ob[: m.conv_out.bias.numel()] = m.conv_out.bias.detach().float()
        self.conv_out_b = ob
        self.temb_w = torch.cat(temb_rows, 0).contiguous()
        self.temb_b = torch.cat(temb_bias, 0).contiguous()
        edim = self.te_w1.shape[0]
        self.te_scratch = torch.empty(2 * edim + self.te_w1.shape[1], dtype=F32, device=dev)
        self.temb_vec = torch.empty(self.temb_w.shape[0], dtype=F32, device=dev)
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    # ------------------------------------------------------------------------------------------
    _ctx_kv: dict = {}
    _ctx_key = None

    def set_process_group(self, pg, rank: int, world: int):
        self.pg, self.rank, self.world = pg, rank, world

    # ------------------------------------------------------------------------------------------
    # blocks
    # ------------------------------------------------------------------------------------------
    def _resnet(self, r, x0, x1, N, H, W):
        hw = H * W
        h = ops.groupnorm(x0, r.n1w, r.n1b, nimg=N, hw=hw, groups=self.groups, eps=self.eps, silu=True,
                          x1=x1)
        h = ops.gemm(h, r.w1, bias=self.temb_vec[r.temb_off:r.temb_off + r.cout], conv=(N, H, W))
        h = ops.groupnorm(h, r.n2w, r.n2b, nimg=N, hw=hw, groups=self.groups, eps=self.eps, silu=True)
        if r.ws is not None:
            sc = ops.gemm(x0, r.ws, a1=x1, bias=r.bs)
        else:
            assert x1 is None
            sc = x0
        return ops.gemm(h, r.w2, bias=r.b2, residual=sc, conv=(N, H, W))

    def _project_ctx_block(self, s, ctx2d, nctx, lctx):
        lcp = (lctx + 7) // 8 * 8
        k2 = torch.empty((nctx * lctx, s.c), dtype=F16, device=self.dev)
        vt2 = torch.empty((nctx, s.c, lcp), dtype=F16, device=self.dev)
        ops.gemm(ctx2d, s.wkv2, outs=[k2, vt2], trans=[False, True], trans_rows=lctx)
        return k2, vt2

    def spatial_entries(self):
        out = []
        for e in list(self.down) + list(self.up):
            out.extend(e.att or [])
        out.append(self.mid_att)
        return out

    def project_context(self, ctx: Optional[torch.Tensor]) -> None:
        """K / V^T of the CLIP context [nctx, L, D] for every spatial block, once (the 257-token
        context is constant over the denoising loop).  run() uses them whenever it is called with this very tensor
        (same storage and shape); the caller re-projects after changing its contents.  None drops the cache."""
        self._ctx_kv, self._ctx_key = {}, None
        if ctx is None:
            return
        nctx, lctx, dctx = ctx.shape
        ctx2d = ctx.reshape(nctx * lctx, dctx)
        for s in self.spatial_entries():
            self._ctx_kv[s.name] = self._project_ctx_block(s, ctx2d, nctx, lctx)
        self._ctx_key = (ctx2d.data_ptr(), nctx, lctx)

    def _ff(self, o, h, lnw, lnb):
        n = ops.layernorm(h, lnw, lnb)
        g = ops.gemm(n, o.ff1_w, bias=o.ff1_b, geglu=True)
        return ops.gemm(g, o.ff2_w, bias=o.ff2_b, residual=h)

    def _spatial(self, s, x, N, H, W, f, ctx2d, nctx, lctx, bank, n_uncond, capture=None):
        """`capture` (reference UNet, write mode): dict that receives norm1's output under s.name."""
        hw = H * W
        C = s.c
        d = C // self.heads
        dev = self.dev
        h = ops.groupnorm(x, s.gnw, s.gnb, nimg=N, hw=hw, groups=self.groups, eps=1e-6, silu=False)
        h = ops.gemm(h, s.pin_w, bias=s.pin_b)
        lp = (hw + 7) // 8 * 8
        # V^T per head padded to a multiple of 16 rows when d % 16 == 8 (d = 40): row d of every head is
        # a row of ones, so the attention kernel's P.V MMA also yields the softmax row sums
        ones = (d % 16 == 8) and d <= 64
        dp = d + 8 if ones else d
        th = (d, dp) if ones else None
        q = torch.empty((N * hw, C), dtype=F16, device=dev)
        k = torch.empty((N * hw, C), dtype=F16, device=dev)
        vt = torch.empty((N, self.heads * dp, lp), dtype=F16, device=dev)
        if ones:
            vt.view(N, self.heads, dp, lp)[:, :, d:].fill_(1.0)
        if bank is not None:
            r0 = n_uncond * hw
            # banks hold the rows that are read: the cond images only (the CFG uncond half never sees the bank,
            # mutual_mix_attention.py:181-201); a full [(N hw), C] bank (the reference's layout) is sliced here
            bank_c = bank if bank.shape[0] == (N - n_uncond) * hw else bank[r0:]
            n1, kvin = ops.layernorm(h, s.ln1w, s.ln1b, add=bank_c, add_row0=r0)
            if n_uncond > 0:     # CFG uncond half: plain self-attention (mutual_mix_attention.py:181-201)
                ops.gemm(n1[:r0], s.wqkv, outs=[q[:r0], k[:r0], vt[:n_uncond]],
                         trans=[False, False, True], trans_rows=hw, trans_head=th)
            ops.gemm(n1[r0:], s.wq, out=q[r0:])
            ops.gemm(kvin, s.wkv, outs=[k[r0:], vt[n_uncond:]], trans=[False, True], trans_rows=hw,
                     trans_head=th)
        else:
            n1 = ops.layernorm(h, s.ln1w, s.ln1b)
            if capture is not None:
                capture[s.name] = n1
            ops.gemm(n1, s.wqkv, outs=[q, k, vt], trans=[False, False, True], trans_rows=hw, trans_head=th)
        a = ops.attention(q, k, vt, nimg=N, lq=hw, lkv=hw, heads=self.heads, d=d, vt_head_rows=dp,
                          vt_ones=ones)
        h = ops.gemm(a, s.wo, bias=s.bo, residual=h)
        # CLIP cross-attention (mutual_mix_attention.py:206-220); K / V^T of the context: projected once per clip
        # by project_context() (the denoising loop's context is constant), else per call
        n2 = ops.layernorm(h, s.ln2w, s.ln2b)
        q2 = ops.gemm(n2, s.wq2)
        cached = self._ctx_kv.get(s.name) if self._ctx_key == (ctx2d.data_ptr(), nctx, lctx) else None
        if cached is not None:
            k2, vt2 = cached
        else:
            k2, vt2 = self._project_ctx_block(s, ctx2d, nctx, lctx)
        a2 = ops.attention(q2, k2, vt2, nimg=N, lq=hw, lkv=lctx, heads=self.heads, d=d,
                           kv_div=max(1, N // nctx))
        h = ops.gemm(a2, s.wo2, bias=s.bo2, residual=h)
        h = self._ff(s, h, s.ln3w, s.ln3b)
        return ops.gemm(h, s.pout_w, bias=s.pout_b, residual=x)

    def _motion_a2a(self, o, x, N, H, W, nb, fl, f_total):
        """Frame-sharded motion module with two all-to-all exchanges (frame-sharded <-> pixel-sharded) per CFG
        branch instead of an all-gather of K/V per temporal attention: everything between GroupNorm (per image)
        and proj_out (per token) is per pixel, so once each rank holds ALL frames of hw/G pixels the temporal
        transformer needs no further communication.  Volume per module: 2 * (G-1)/G of the local activation
        (4-16x less than the K/V all-gather).  No layout copies around the exchanges: GroupNorm
        writes the send buffer ([G(dst), fl, pp, C], `chunks=`), the received blocks already are frame-major
        ([F, pp, C] per branch, also with an uneven frame split), and one `unshard` kernel per branch puts the
        returning pixel chunks side by side (pure data movement: the sharded result equals the single-GPU one
        bit for bit)."""
        from .sharding import frame_split, frames_to_pixels, pixels_per_rank, pixels_to_frames
        hw, C, G = H * W, o.c, self.world
        counts = frame_split(f_total, G)
        assert counts[self.rank] == fl, (counts, self.rank, fl)
        pp = pixels_per_rank(hw, G)
        d = C // self.mheads
        recv = torch.empty((nb, f_total * pp, C), dtype=F16, device=self.dev)
        for b in range(nb):
            send = ops.groupnorm(x[b * fl * hw:(b + 1) * fl * hw], o.gnw, o.gnb, nimg=fl, hw=hw, groups=self.groups,
                                 eps=1e-6, silu=False, chunks=(G, pp))
            frames_to_pixels(send, recv[b], counts, pp, self.rank, self.pg)
        h = ops.gemm(recv.view(nb * f_total * pp, C), o.pin_w, bias=o.pin_b)    # rows [(b f_total) pp]
        for e in o.att:
            n = ops.layernorm(h, e.lnw, e.lnb)
            qkv = ops.gemm(n, e.wqkv, row_bias=e.pe_qkv[:f_total], row_div=pp)
            a = ops.temporal_attention(qkv, nb=nb, f_q=f_total, npix=pp, heads=self.mheads, d=d)
            h = ops.gemm(a, e.wo, bias=e.bo, residual=h)
        h = self._ff(o, h, o.ffnw, o.ffnb)
        hb = torch.empty((nb * fl * hw, C), dtype=F16, device=self.dev)
        for b in range(nb):
            back = torch.empty((G * fl * pp, C), dtype=F16, device=self.dev)
            pixels_to_frames(h[b * f_total * pp:(b + 1) * f_total * pp], back, counts, pp, self.rank, self.pg)
            ops.unshard(back, nimg=fl, hw=hw, chunk_pix=pp, out=hb[b * fl * hw:(b + 1) * fl * hw])
        return ops.gemm(hb, o.pout_w, bias=o.pout_b, residual=x)

    def _motion(self, o, x, N, H, W, nb, fl, f_off, f_total):
        hw = H * W
        if self.world > 1 and self.shard_mode == "a2a":
            return self._motion_a2a(o, x, N, H, W, nb, fl, f_total)
        C = o.c
        d = C // self.mheads
        h = ops.groupnorm(x, o.gnw, o.gnb, nimg=N, hw=hw, groups=self.groups, eps=1e-6, silu=False)
        h = ops.gemm(h, o.pin_w, bias=o.pin_b)
        for e in o.att:
            n = ops.layernorm(h, e.lnw, e.lnb)
            pe_rows = e.pe_qkv[f_off:f_off + fl]        # frame j of this shard sits at window position f_off+j
            if self.world == 1 and not self.force_simt:
                qkv = ops.gemm(n, e.wqkv, row_bias=pe_rows, row_div=hw)
                a = ops.temporal_attention(qkv, nb=nb, f_q=fl, npix=hw, heads=self.mheads, d=d)
            else:
                import torch.distributed as dist
                qb = torch.empty((N * hw, C), dtype=F16, device=self.dev)
                kv = torch.empty((N * hw, 2 * C), dtype=F16, device=self.dev)
                ops.gemm(n, e.wqkv, outs=[qb, kv[:, :C], kv[:, C:]], row_bias=pe_rows, row_div=hw)
                if self.world > 1:
                    if f_total != fl * self.world:
                        raise NotImplementedError("the K/V all-gather mode needs an even frame split "
                                                  "(use shard_mode='a2a')")
                    kv_all = torch.empty((self.world * N * hw, 2 * C), dtype=F16, device=self.dev)
                    dist.all_gather_into_tensor(kv_all, kv, group=self.pg)   # NCCL over NVLink
                else:
                    kv_all = kv      # force_simt: the gathered-layout kernel on one GPU (multi-GPU parity anchor)
                a = ops.temporal_attention(qb, nb=nb, f_q=fl, npix=hw, heads=self.mheads, d=d,
                                           kv=kv_all, f_kv=f_total, f_kv_rank=fl,
                                           f_q_offset=f_off, kv_offsets=(0, C))
            h = ops.gemm(a, e.wo, bias=e.bo, residual=h)
        h = self._ff(o, h, o.ffnw, o.ffnb)
        return ops.gemm(h, o.pout_w, bias=o.pout_b, residual=x)

    # ------------------------------------------------------------------------------------------
    def run(self, x_in: torch.Tensor, nb: int, fl: int, H: int, W: int, ctx: torch.Tensor,
            banks: Optional[Dict[str, torch.Tensor]], n_uncond: int, f_off: int = 0,
            f_total: Optional[int] = None) -> torch.Tensor:
        """x_in: [(nb fl) H W, cin_pad] fp16 NHWC latents; the timestep is read from self.t_dev;
        ctx [nctx, L, D] fp16; banks {module path: [(nb fl) hw, C] fp16}.  Returns [(nb fl) H W, 8]
        fp16 (columns >= out_channels are zero)."""
        N = nb * fl
        f_total = f_total if f_total is not None else fl
        cfg = self.cfg
        ops.time_embed(self.t_dev, self.te_w1, self.te_b1, self.te_w2, self.te_b2, self.temb_w,
                       self.temb_b, flip_sin_to_cos=bool(cfg["flip_sin_to_cos"]),
                       freq_shift=float(cfg["freq_shift"]), scratch=self.te_scratch, out=self.temb_vec)
        nctx, lctx, dctx = ctx.shape
        ctx2d = ctx.reshape(nctx * lctx, dctx)
        banks = banks or {}

        def bank_of(s):
            b = banks.get(s.name)
            if b is None:
                return None
            return b.reshape(-1, b.shape[-1])

        def tr(key, t, hh, ww):
            if self.trace is not None:
                self.trace[key] = (t.clone(), N, hh, ww)

        x = ops.gemm(x_in, self.conv_in_w, bias=self.conv_in_b, conv=(N, H, W))
        tr("conv_in", x, H, W)
        skips = [(x, H, W)]
        h_, w_ = H, W
        for e in self.down:
            for j, r in enumerate(e.res):
                x = self._resnet(r, x, None, N, h_, w_)
                if e.att is not None:
                    s = e.att[j]
                    x = self._spatial(s, x, N, h_, w_, fl, ctx2d, nctx, lctx, bank_of(s), n_uncond)
                x = self._motion(e.mm[j], x, N, h_, w_, nb, fl, f_off, f_total)
                tr(f"down_blocks.{e.plan['idx']}.{j}", x, h_, w_)
                skips.append((x, h_, w_))
            if e.plan["downsample"]:
                col = ops.im2col3x3(x, N, h_, w_, 2)
                h_, w_ = (h_ - 1) // 2 + 1, (w_ - 1) // 2 + 1
                x = ops.gemm(col, e.ds_w, bias=e.ds_b)
                skips.append((x, h_, w_))
        x = self._resnet(self.mid_res[0], x, None, N, h_, w_)
        s = self.mid_att
        x = self._spatial(s, x, N, h_, w_, fl, ctx2d, nctx, lctx, bank_of(s), n_uncond)
        x = self._motion(self.mid_mm, x, N, h_, w_, nb, fl, f_off, f_total)
        x = self._resnet(self.mid_res[1], x, None, N, h_, w_)
        tr("mid", x, h_, w_)
        for e in self.up:
            for j, r in enumerate(e.res):
                skip, sh, sw = skips.pop()
                assert (sh, sw) == (h_, w_), "latent size must be divisible by 8"
                x = self._resnet(r, x, skip, N, h_, w_)
                if e.att is not None:
                    s = e.att[j]
                    x = self._spatial(s, x, N, h_, w_, fl, ctx2d, nctx, lctx, bank_of(s), n_uncond)
                x = self._motion(e.mm[j], x, N, h_, w_, nb, fl, f_off, f_total)
                tr(f"up_blocks.{e.plan['idx']}.{j}", x, h_, w_)
            if e.plan["upsample"]:
                x = ops.upsample2x(x, N, h_, w_)
                h_, w_ = 2 * h_, 2 * w_
                x = ops.gemm(x, e.us_w, bias=e.us_b, conv=(N, h_, w_))
        x = ops.groupnorm(x, self.out_nw, self.out_nb, nimg=N, hw=h_ * w_, groups=self.groups,
                          eps=self.eps, silu=True)
        return ops.gemm(x, self.conv_out_w, bias=self.conv_out_b, conv=(N, h_, w_))

    # ------------------------------------------------------------------------------------------
    def set_timestep(self, timestep) -> None:
        if torch.is_tensor(timestep):
            t = timestep.reshape(-1)
            if t.numel() > 1 and not bool((t == t[0]).all()):
                raise NotImplementedError("per-sample timesteps are not supported (the reference's "
                                          "denoising loop passes one scalar timestep)")
            self.t_dev.copy_(t[:1].to(device=self.dev, dtype=torch.int64))
        else:
            self.t_dev.fill_(int(timestep))

    def collect_banks(self, N: int):
        """Banks installed on the blocks by ReferenceAttentionControl (read mode)."""
        banks = {}
        ctrl = self.model._ref_control
        if ctrl is None or ctrl.get("mode") != "read":
            return banks, 0
        for e in list(self.down) + list(self.up):
            for s in (e.att or []):
                if len(s.block.bank) == 1:
                    banks[s.name] = s.block.bank[0]
        if len(self.mid_att.block.bank) == 1:
            banks[self.mid_att.name] = self.mid_att.block.bank[0]
        n_uncond = N // 2 if ctrl.get("do_classifier_free_guidance") else 0
        return banks, n_uncond

    def forward_api(self, sample: torch.Tensor, timestep, ctx: torch.Tensor) -> torch.Tensor:
        """The UNet3DConditionModel.forward call surface: sample [B, 4, f, h, w] -> same shape."""
        B, Cc, f, H, W = sample.shape
        if H % 8 or W % 8:
            raise ValueError(f"latent size {H}x{W} must be divisible by 8")
        if f > self.cfg["pe_max_len"]:
            raise ValueError(f"window of {f} frames exceeds temporal_position_encoding_max_len="
                             f"{self.cfg['pe_max_len']}")
        dev = self.dev
        self.set_timestep(timestep)
        s16 = sample.to(device=dev, dtype=F16).contiguous()
        x_in = ops.latents_to_nhwc(s16, b=B, frame_idx=None, fl=f, cpad=self.cin_pad)
        c16 = ctx.to(device=dev, dtype=F16).contiguous()
        if c16.shape[0] not in (1, B):
            raise ValueError("encoder_hidden_states batch must be 1 or the sample batch")
        banks, n_uncond = self.collect_banks(B * f)
        banks = {k: v.to(device=dev, dtype=F16).contiguous() for k, v in banks.items()}
        y = self.run(x_in, B, f, H, W, c16, banks, n_uncond)
        co = self.cfg["out_channels"]
        out = y[:, :co].reshape(B, f, H, W, co).permute(0, 4, 1, 2, 3).contiguous()
        return out.to(sample.dtype)
