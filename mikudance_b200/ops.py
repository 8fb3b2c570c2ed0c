"""Tensor-level wrappers over the C ABI (include/mdk.h).  PyTorch tensors are only containers for
device memory here; all arithmetic happens in the sm_100a kernels of libmikudance_sm100.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from ._lib import (AttnArgs, GemmArgs, GnArgs, LnArgs, ManArgs, TattnArgs, TembArgs, check, cur_stream,
                   get_ctx, load_library, ptr)

F16 = torch.float16


class Profiler:
    """CUDA-event timing of every kernel launch issued through this module (bench.py's roofline
    pass; never active inside a timed region or a graph capture)."""

    def __init__(self):
        self.records = []   # (kind, flops, bytes, ev0, ev1)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(kind.split(":")[0], dict(n=0, ms=0.0, flops=0.0, bytes=0.0))
            d["n"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
            d["bytes"] += nbytes
        for d in out.values():
            if d["ms"] > 0:
                d["tflops"] = d["flops"] / (d["ms"] * 1e-3) / 1e12
                d["gbs"] = d["bytes"] / (d["ms"] * 1e-3) / 1e9
        return out


    def shapes(self, top=16):
        """per-shape breakdown (kind strings carry the problem shape after the first ':')"""
        torch.cuda.synchronize()
        out = {}
        for kind, flops, nbytes, e0, e1 in self.records:
            d = out.setdefault(kind, dict(n=0, ms=0.0, flops=0.0))
            d["n"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += flops
        rows = sorted(out.items(), key=lambda kv: -kv[1]["ms"])[:top]
        return [dict(shape=k, n=v["n"], ms=round(v["ms"], 3),
                     tflops=round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] > 0 else 0.0) for k, v in rows]


_PROF: Optional[Profiler] = None


def set_profiler(p: Optional[Profiler]) -> None:
    global _PROF
    _PROF = p


def _run(kind: str, flops: float, nbytes: float, fn, what: str) -> None:
    if _PROF is None:
        check(fn(), what)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(), what)
    e1.record()
    _PROF.records.append((kind, flops, nbytes, e0, e1))


def _chk16(t: torch.Tensor, name: str) -> None:
    if t.dtype != F16 or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA float16 tensor, got {t.dtype} on {t.device}")


def gemm(a0: torch.Tensor, w: torch.Tensor, *, a1: Optional[torch.Tensor] = None,
         bias: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
         row_div: int = 1, residual: Optional[torch.Tensor] = None, geglu: bool = False,
         conv: Optional[Tuple[int, int, int]] = None, out: Optional[torch.Tensor] = None,
         outs: Optional[Sequence[torch.Tensor]] = None, trans: Sequence[bool] = (False, False, False),
         trans_rows: int = 0, trans_head: Optional[Tuple[int, int]] = None) -> torch.Tensor | Sequence[torch.Tensor]:
    """D = [a0 | a1] @ w^T with the fused epilogue of mdk_gemm_f16.

    a0/a1: [M, K_i] (conv=None) or NHWC [nimg, h, w, C_i] flattened to [M, C_i] with conv=(nimg,h,w)
    w: [N, K] fp16;  bias: fp32 [N];  row_bias: fp32 [row_mod, N];  residual: fp16 [M, N_out]
    outs: up to 3 equally wide column segments (each [M, seg] or, if trans[s], [nimg, seg, ld]).
    """
    _chk16(a0, "a0"), _chk16(w, "w")
    dev = a0.device
    lib = load_library()
    M = a0.shape[0]
    N = w.shape[0]
    k0 = a0.shape[1]
    k1 = a1.shape[1] if a1 is not None else 0
    args = GemmArgs()
    args.a0, args.lda0, args.k0 = ptr(a0), a0.stride(0), k0
    if a1 is not None:
        _chk16(a1, "a1")
        assert a1.shape[0] == M
        args.a1, args.lda1, args.k1 = ptr(a1), a1.stride(0), k1
    args.b, args.ldb = ptr(w), w.stride(0)
    args.m, args.n = M, N
    if conv is not None:
        nimg, h, wd = conv
        assert a0.is_contiguous() and (a1 is None or a1.is_contiguous())
        args.conv_taps, args.nimg, args.h, args.w = 9, nimg, h, wd
        assert w.shape[1] == 9 * (k0 + k1), (w.shape, k0, k1)
    else:
        args.conv_taps = 1
        assert w.shape[1] == k0 + k1, (w.shape, k0, k1)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        args.bias = ptr(bias)
    if row_bias is not None:
        assert row_bias.dtype == torch.float32 and row_bias.shape[1] == N and row_bias.is_contiguous()
        args.row_bias, args.row_div, args.row_mod = ptr(row_bias), row_div, row_bias.shape[0]
    n_out = N // 2 if geglu else N
    if residual is not None:
        _chk16(residual, "residual")
        assert residual.shape == (M, n_out)
        args.residual, args.ldr = ptr(residual), residual.stride(0)
    args.geglu = 1 if geglu else 0
    if outs is None:
        if out is None:
            out = torch.empty((M, n_out), dtype=F16, device=dev)
        assert out.shape == (M, n_out)
        args.out[0], args.ldo[0] = ptr(out), out.stride(0)
        ret = out
    else:
        nseg = len(outs)
        assert n_out % nseg == 0
        seg = n_out // nseg
        args.seg_cols = seg
        for s, o in enumerate(outs):
            _chk16(o, f"outs[{s}]")
            args.out[s] = ptr(o)
            args.out_trans[s] = 1 if trans[s] else 0
            if trans[s]:
                if trans_head is not None:
                    hd, hdp = trans_head
                    assert o.dim() == 3 and o.shape[1] == (seg // hd) * hdp, o.shape
                    args.trans_head_d, args.trans_head_dp = hd, hdp
                else:
                    assert o.dim() == 3 and o.shape[1] == seg, o.shape
                args.trans_ld = o.stride(1)
                args.trans_rows = trans_rows
            else:
                args.ldo[s] = o.stride(0)
        ret = outs
    ktot = (9 if conv is not None else 1) * (k0 + k1)
    tag = "gemm_tc"
    if _PROF is not None:
        tag = (f"gemm_tc:{'conv' if conv is not None else 'lin'} M={M} N={N} K={ktot}"
               f"{' bias' if bias is not None else ''}{' rowbias' if row_bias is not None else ''}"
               f"{' res' if residual is not None else ''}{' geglu' if geglu else ''}"
               f"{' seg' + str(len(outs)) if outs is not None else ''}{' T' if (outs is not None and any(trans)) else ''}")
    _run(tag, 2.0 * M * N * ktot, 2.0 * (M * (k0 + k1) + N * ktot + M * n_out * (2 if residual is not None else 1)),
         lambda: lib.mdk_gemm_f16(get_ctx(dev), C.byref(args), cur_stream(dev)), "mdk_gemm_f16")
    return ret


def attention(q: torch.Tensor, k: torch.Tensor, vt: torch.Tensor, *, nimg: int, lq: int, lkv: int,
              heads: int, d: int, kv_div: int = 1, scale: Optional[float] = None,
              out: Optional[torch.Tensor] = None, vt_head_rows: int = 0, vt_ones: bool = False) -> torch.Tensor:
    """q [nimg*lq, heads*d], k [nkv*lkv, heads*d], vt [nkv, heads*d, ldvt] -> out [nimg*lq, heads*d]"""
    _chk16(q, "q"), _chk16(k, "k"), _chk16(vt, "vt")
    dev = q.device
    if out is None:
        out = torch.empty((nimg * lq, heads * d), dtype=F16, device=dev)
    a = AttnArgs()
    a.q, a.k, a.vt, a.out = ptr(q), ptr(k), ptr(vt), ptr(out)
    a.ldq, a.ldk, a.ldvt, a.ldo = q.stride(0), k.stride(0), vt.stride(1), out.stride(0)
    a.nimg, a.nkv, a.kv_div = nimg, vt.shape[0], kv_div
    a.lq, a.lkv, a.heads, a.d = lq, lkv, heads, d
    a.scale = scale if scale is not None else 1.0 / math.sqrt(d)
    a.vt_head_rows, a.vt_ones = vt_head_rows, 1 if vt_ones else 0
    _run(f"attn_tc:L={lq}x{lkv} d={d} n={nimg}" if _PROF is not None else "attn_tc",
         4.0 * nimg * heads * lq * lkv * d, 2.0 * heads * d * (2 * nimg * lq + 2 * vt.shape[0] * lkv),
         lambda: load_library().mdk_attn_fwd_f16(get_ctx(dev), C.byref(a), cur_stream(dev)),
         "mdk_attn_fwd_f16")
    return out


def temporal_attention(qkv: torch.Tensor, *, nb: int, f_q: int, npix: int, heads: int, d: int,
                       pe_q: Optional[torch.Tensor] = None, kv: Optional[torch.Tensor] = None,
                       f_kv: Optional[int] = None, f_kv_rank: int = 0, f_q_offset: int = 0,
                       kv_offsets: Optional[Tuple[int, int]] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv [(nb f_q) npix, 3C] fused q|k|v rows; kv (optional) gathered K/V rows of all frames."""
    _chk16(qkv, "qkv")
    dev = qkv.device
    Cc = heads * d
    if out is None:
        out = torch.empty((nb * f_q * npix, Cc), dtype=F16, device=dev)
    a = TattnArgs()
    a.q, a.q_ld, a.q_off = ptr(qkv), qkv.stride(0), 0
    if kv is None:
        a.kv, a.kv_ld, a.k_off, a.v_off = ptr(qkv), qkv.stride(0), Cc, 2 * Cc
        a.f_kv = f_q
    else:
        _chk16(kv, "kv")
        ko, vo = kv_offsets if kv_offsets is not None else (0, Cc)
        a.kv, a.kv_ld, a.k_off, a.v_off = ptr(kv), kv.stride(0), ko, vo
        a.f_kv = f_kv
    a.f_kv_rank = f_kv_rank
    if pe_q is not None:
        assert pe_q.dtype == torch.float32 and pe_q.shape[1] == Cc and pe_q.is_contiguous()
        a.pe_q = ptr(pe_q)
    a.out, a.out_ld = ptr(out), out.stride(0)
    a.nb, a.f_q, a.f_q_offset, a.npix, a.heads, a.d = nb, f_q, f_q_offset, npix, heads, d
    a.scale = 1.0 / math.sqrt(d)
    fkv = a.f_kv
    _run("temporal_attn", 4.0 * nb * npix * heads * f_q * fkv * d, 2.0 * nb * npix * Cc * (2 * f_q + 2 * fkv),
         lambda: load_library().mdk_temporal_attn_f16(get_ctx(dev), C.byref(a), cur_stream(dev)),
         "mdk_temporal_attn_f16")
    return out


def groupnorm(x0: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, nimg: int, hw: int,
              groups: int, eps: float, silu: bool, x1: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None,
              chunks: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """chunks = (n_chunks, chunk_pix): write the exchange layout [n_chunks, nimg, chunk_pix, C] (the send buffer of
    the frames -> pixels all-to-all; rows of pixels >= hw stay zero) instead of [nimg * hw, C]."""
    _chk16(x0, "x0")
    dev = x0.device
    c0 = x0.shape[1]
    c1 = x1.shape[1] if x1 is not None else 0
    if out is None:
        if chunks is None:
            out = torch.empty((nimg * hw, c0 + c1), dtype=F16, device=dev)
        elif chunks[0] * chunks[1] == hw:
            out = torch.empty((chunks[0] * nimg * chunks[1], c0 + c1), dtype=F16, device=dev)
        else:
            out = torch.zeros((chunks[0] * nimg * chunks[1], c0 + c1), dtype=F16, device=dev)
    if ws is None:
        ws = torch.empty((load_library().mdk_groupnorm_ws_bytes(nimg, groups),), dtype=torch.uint8, device=dev)
    a = GnArgs()
    a.x0, a.c0 = ptr(x0), c0
    if x1 is not None:
        a.x1, a.c1 = ptr(x1), c1
    a.nimg, a.hw, a.groups, a.eps = nimg, hw, groups, eps
    a.gamma, a.beta, a.silu = ptr(gamma), ptr(beta), 1 if silu else 0
    a.out, a.ws = ptr(out), ptr(ws)
    if chunks is not None:
        a.out_chunks, a.out_chunk_pix = chunks
    _run("groupnorm", 0.0, 3.0 * 2 * nimg * hw * (c0 + c1),
         lambda: load_library().mdk_groupnorm_f16(get_ctx(dev), C.byref(a), cur_stream(dev)),
         "mdk_groupnorm_f16")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, eps: float = 1e-5,
              add: Optional[torch.Tensor] = None, add_row0: int = 0,
              out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None):
    _chk16(x, "x")
    dev = x.device
    rows, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    a = LnArgs()
    a.x, a.rows, a.c, a.eps = ptr(x), rows, c, eps
    a.gamma, a.beta, a.out = ptr(gamma), ptr(beta), ptr(out)
    if add is not None:
        _chk16(add, "add")
        assert add.shape == (rows - add_row0, c) and add.is_contiguous()
        if out2 is None:
            out2 = torch.empty_like(add)
        a.add, a.out2, a.add_row0 = ptr(add), ptr(out2), add_row0
    _run("layernorm", 0.0, 2.0 * rows * c * (2 if add is None else 3),
         lambda: load_library().mdk_layernorm_f16(get_ctx(dev), C.byref(a), cur_stream(dev)),
         "mdk_layernorm_f16")
    return (out, out2) if add is not None else out


def upsample2x(x: torch.Tensor, nimg: int, h: int, w: int) -> torch.Tensor:
    _chk16(x, "x")
    c = x.shape[1]
    out = torch.empty((nimg * 4 * h * w, c), dtype=F16, device=x.device)
    _run("upsample2x", 0.0, 2.0 * 5 * nimg * h * w * c,
         lambda: load_library().mdk_upsample2x_f16(get_ctx(x.device), ptr(x), ptr(out), nimg, h, w, c,
                                                   cur_stream(x.device)), "mdk_upsample2x_f16")
    return out


def im2col3x3(x: torch.Tensor, nimg: int, h: int, w: int, stride: int) -> torch.Tensor:
    _chk16(x, "x")
    c = x.shape[1]
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    kpad = 9 * c
    out = torch.empty((nimg * ho * wo, kpad), dtype=F16, device=x.device)
    _run("im2col", 0.0, 2.0 * (nimg * h * w * c + out.numel()),
         lambda: load_library().mdk_im2col3x3_f16(get_ctx(x.device), ptr(x), ptr(out), nimg, h, w, c,
                                                  stride, kpad, cur_stream(x.device)), "mdk_im2col3x3_f16")
    return out


def time_embed(timestep: torch.Tensor, w1, b1, w2, b2, proj_w, proj_b, *, flip_sin_to_cos: bool,
               freq_shift: float, scratch: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    dev = w1.device
    assert timestep.dtype == torch.int64 and timestep.is_cuda
    a = TembArgs()
    a.timestep, a.dim = ptr(timestep), w1.shape[1]
    a.flip_sin_to_cos, a.freq_shift = 1 if flip_sin_to_cos else 0, float(freq_shift)
    a.w1, a.b1, a.w2, a.b2, a.edim = ptr(w1), ptr(b1), ptr(w2), ptr(b2), w1.shape[0]
    a.proj_w, a.proj_b, a.nrows = ptr(proj_w), ptr(proj_b), (proj_w.shape[0] if proj_w is not None else 0)
    a.scratch, a.temb_out = ptr(scratch), ptr(out)
    _run("time_embed", 0.0, 2.0 * (w1.numel() + w2.numel() + (proj_w.numel() if proj_w is not None else 0)),
         lambda: load_library().mdk_time_embed_f16(get_ctx(dev), C.byref(a), cur_stream(dev)),
         "mdk_time_embed_f16")
    return out


def latents_to_nhwc(sample: torch.Tensor, *, b: int, frame_idx: Optional[torch.Tensor], fl: int,
                    cpad: int) -> torch.Tensor:
    """sample [b_src, c, F, h, w] fp16 -> [(b fl) h w, cpad]"""
    _chk16(sample, "sample")
    assert sample.is_contiguous()
    b_src, c, F, h, w = sample.shape
    out = torch.empty((b * fl * h * w, cpad), dtype=F16, device=sample.device)
    _run("latent_glue", 0.0, 2.0 * out.numel(),
         lambda: load_library().mdk_latents_to_nhwc(get_ctx(sample.device), ptr(sample), ptr(out), b,
                                                    b_src, c, F, ptr(frame_idx), fl, h * w, cpad,
                                                    cur_stream(sample.device)), "mdk_latents_to_nhwc")
    return out


def pred_accumulate(pred: torch.Tensor, acc: torch.Tensor, counter: Optional[torch.Tensor], *,
                    frame_idx: Optional[torch.Tensor], fl: int) -> None:
    """pred [(b fl) hw, cpad] fp16;  acc [b, c, F, h, w] fp32 +=;  counter [F] fp32 += 1"""
    b, c, F, h, w = acc.shape
    assert acc.dtype == torch.float32 and acc.is_contiguous()
    _run("latent_glue", 0.0, 2.0 * pred.numel() + 8.0 * b * c * fl * h * w,
         lambda: load_library().mdk_pred_accumulate(get_ctx(pred.device), ptr(pred), ptr(acc),
                                                    ptr(counter), b, c, F, ptr(frame_idx), fl, h * w,
                                                    pred.shape[1], cur_stream(pred.device)),
         "mdk_pred_accumulate")


def unshard(back: torch.Tensor, *, nimg: int, hw: int, chunk_pix: int, x: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i, px] = back[px // chunk_pix, i, px % chunk_pix] (+ x[i, px]): the way back from the pixel-sharded layout
    of a frame-sharded motion module ([chunks, nimg, chunk_pix, C] -> [nimg * hw, C]), optionally fused with a
    residual add."""
    _chk16(back, "back")
    c = back.shape[1]
    assert back.shape[0] >= nimg * hw and back.is_contiguous()
    if x is not None:
        _chk16(x, "x")
        assert x.shape == (nimg * hw, c) and x.is_contiguous()
    if out is None:
        out = torch.empty((nimg * hw, c), dtype=F16, device=back.device)
    _run("latent_glue", 0.0, (6.0 if x is not None else 4.0) * out.numel(),
         lambda: load_library().mdk_unshard_add_f16(get_ctx(back.device), ptr(back), ptr(x), ptr(out), nimg, hw,
                                                    chunk_pix, c, cur_stream(back.device)), "mdk_unshard_add_f16")
    return out


def cfg_ddim_step(acc: torch.Tensor, counter: torch.Tensor, latents: torch.Tensor,
                  coef: torch.Tensor, guidance_scale: float, v_prediction: bool) -> None:
    """in-place DDIM update of latents [1, c, F, h, w] fp16 from acc [nb, c, F, h, w] fp32"""
    nb, c, F, h, w = acc.shape
    _chk16(latents, "latents")
    assert latents.is_contiguous() and coef.dtype == torch.float32 and coef.numel() == 4
    _run("cfg_ddim", 0.0, 4.0 * acc.numel() + 4.0 * latents.numel(),
         lambda: load_library().mdk_cfg_ddim_step(get_ctx(acc.device), ptr(acc), ptr(counter),
                                                  ptr(latents), ptr(coef), float(guidance_scale), nb, c,
                                                  F, h * w, 1 if v_prediction else 0,
                                                  cur_stream(acc.device)), "mdk_cfg_ddim_step")


# ------------------------------------------------------------------------------------------------
# reference UNet (writer) only — SURVEY.md §8f row 1
# ------------------------------------------------------------------------------------------------
def cond_to_nhwc(x: torch.Tensor, *, c_first: int, c: int, ho: int, wo: int, cpad: int) -> torch.Tensor:
    """x [nimg, ctot, h, w] fp16 NCHW -> [(nimg ho wo), cpad] fp16: channels [c_first, c_first+c),
    nearest-resized to (ho, wo), zero-padded to cpad columns."""
    _chk16(x, "x")
    assert x.dim() == 4 and x.is_contiguous()
    nimg, ctot, h, w = x.shape
    out = torch.empty((nimg * ho * wo, cpad), dtype=F16, device=x.device)
    _run("cond_glue", 0.0, 2.0 * (nimg * h * w * c + out.numel()),
         lambda: load_library().mdk_cond_to_nhwc_f16(get_ctx(x.device), ptr(x), ptr(out), nimg, ctot,
                                                     c_first, c, h, w, ho, wo, cpad,
                                                     cur_stream(x.device)), "mdk_cond_to_nhwc_f16")
    return out


def relu_(x: torch.Tensor) -> torch.Tensor:
    _chk16(x, "x")
    assert x.is_contiguous()
    _run("relu", 0.0, 4.0 * x.numel(),
         lambda: load_library().mdk_relu_f16(get_ctx(x.device), ptr(x), x.numel(), cur_stream(x.device)),
         "mdk_relu_f16")
    return x


def man_modulate(x: torch.Tensor, gb: torch.Tensor, *, nimg: int, hw: int, eps: float = 1e-5) -> torch.Tensor:
    """x [(nimg hw), C]; gb [(nimg hw), 2C] = gamma | beta -> InstanceNorm(x) * (1 + gamma) + beta"""
    _chk16(x, "x"), _chk16(gb, "gb")
    c = x.shape[1]
    assert x.is_contiguous() and x.shape[0] == nimg * hw and gb.shape == (nimg * hw, 2 * c)
    dev = x.device
    out = torch.empty_like(x)
    ws = torch.empty((load_library().mdk_man_ws_bytes(nimg, c),), dtype=torch.uint8, device=dev)
    a = ManArgs()
    a.x, a.gb, a.ldgb = ptr(x), ptr(gb), gb.stride(0)
    a.nimg, a.hw, a.c, a.eps = nimg, hw, c, eps
    a.out, a.ws = ptr(out), ptr(ws)
    _run("man_modulate", 0.0, 10.0 * nimg * hw * c,
         lambda: load_library().mdk_man_modulate_f16(get_ctx(dev), C.byref(a), cur_stream(dev)),
         "mdk_man_modulate_f16")
    return out


# ------------------------------------------------------------------------------------------------
# CLIP image encoder only — SURVEY.md §8f row 3
# ------------------------------------------------------------------------------------------------
def quick_gelu_(x: torch.Tensor) -> torch.Tensor:
    _chk16(x, "x")
    assert x.is_contiguous()
    _run("quick_gelu", 0.0, 4.0 * x.numel(),
         lambda: load_library().mdk_quick_gelu_f16(get_ctx(x.device), ptr(x), x.numel(), cur_stream(x.device)),
         "mdk_quick_gelu_f16")
    return x


# ------------------------------------------------------------------------------------------------
# VAE only — SURVEY.md §8f row 2
# ------------------------------------------------------------------------------------------------
def softmax_rows_(x: torch.Tensor) -> torch.Tensor:
    """in-place softmax over the columns of x [rows, cols] fp16 (row stride x.stride(0))"""
    _chk16(x, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    _run("softmax_rows", 0.0, 8.0 * rows * cols,
         lambda: load_library().mdk_softmax_rows_f16(get_ctx(x.device), ptr(x), rows, cols, x.stride(0),
                                                     cur_stream(x.device)), "mdk_softmax_rows_f16")
    return x


def im2col3x3_ex(x: torch.Tensor, nimg: int, h: int, w: int, stride: int, pad_lo: int) -> torch.Tensor:
    _chk16(x, "x")
    c = x.shape[1]
    ho, wo = (h + pad_lo - 2) // stride + 1, (w + pad_lo - 2) // stride + 1
    out = torch.empty((nimg * ho * wo, 9 * c), dtype=F16, device=x.device)
    _run("im2col", 0.0, 2.0 * (nimg * h * w * c + out.numel()),
         lambda: load_library().mdk_im2col3x3_ex_f16(get_ctx(x.device), ptr(x), ptr(out), nimg, h, w, c, stride,
                                                     pad_lo, 9 * c, cur_stream(x.device)), "mdk_im2col3x3_ex_f16")
    return out
