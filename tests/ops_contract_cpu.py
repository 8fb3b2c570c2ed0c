"""TEST INFRASTRUCTURE — a CPU statement of the CONTRACT of every `mikudance_b200.ops` wrapper
(= the C ABI in include/mdk.h), in plain PyTorch: float64 math, one rounding to fp16 per output (float64
so that results do not depend on how rows are split over processes or on the BLAS blocking: a sharded run
must then equal the single-process run bit for bit, as it does on the GPU).

Purpose: run the HOST orchestration of the product (`UNetEngine.run`, `RefUNetEngine.run`: weight
packing, NHWC / V^T / GEGLU-panel layouts, skip concat, bank capture, MAN wiring, PE row bias, …)
against the oracle on CPU, where no GPU is available.  The tests monkeypatch `mikudance_b200.ops`
with these functions (`install(monkeypatch)`); nothing in the product imports this module, and the
product's own wrappers refuse non-CUDA tensors.  The CUDA kernels themselves are checked against the
oracle by the `-m gpu` tests.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

F16 = torch.float16


def _h(x):
    return x.to(F16)


def gemm(a0, w, *, a1=None, bias=None, row_bias=None, row_div=1, residual=None, geglu=False, conv=None,
         out=None, outs=None, trans=(False, False, False), trans_rows=0, trans_head=None):
    A = a0.double() if a1 is None else torch.cat([a0.double(), a1.double()], 1)
    M, ktap = A.shape
    N = w.shape[0]
    if conv is not None:
        nimg, h, wd = conv
        assert M == nimg * h * wd and w.shape[1] == 9 * ktap
        x = A.view(nimg, h, wd, ktap).permute(0, 3, 1, 2)
        wt = w.double().view(N, 3, 3, ktap).permute(0, 3, 1, 2)          # K ordered (kh, kw, c)
        D = F.conv2d(x, wt, padding=1).permute(0, 2, 3, 1).reshape(M, N)
    else:
        assert w.shape[1] == ktap
        D = A @ w.double().t()
    if bias is not None:
        assert bias.dtype == torch.float32
        D = D + bias[None, :]
    if row_bias is not None:
        rows = (torch.arange(M) // row_div) % row_bias.shape[0]
        D = D + row_bias[rows]
    if geglu:
        blk = 256
        assert N % blk == 0
        D = D.view(M, N // blk, 2, blk // 2)
        D = (D[:, :, 0] * F.gelu(D[:, :, 1])).reshape(M, N // 2)
    if residual is not None:
        D = D + residual.double()
    D = _h(D)
    if outs is None:
        if out is None:
            return D
        out.copy_(D)
        return out
    nseg = len(outs)
    seg = D.shape[1] // nseg
    for s, o in enumerate(outs):
        Ds = D[:, s * seg:(s + 1) * seg]
        if not trans[s]:
            o.copy_(Ds)
            continue
        nim = M // trans_rows
        T = Ds.view(nim, trans_rows, seg).permute(0, 2, 1)               # [img, col, row-in-image]
        if trans_head is not None:
            hd, hdp = trans_head
            o4 = o.view(nim, seg // hd, hdp, o.shape[2])
            o4[:, :, :hd, :trans_rows] = T.reshape(nim, seg // hd, hd, trans_rows)
        else:
            o[:, :, :trans_rows] = T
    return outs


def attention(q, k, vt, *, nimg, lq, lkv, heads, d, kv_div=1, scale=None, out=None, vt_head_rows=0,
              vt_ones=False):
    scale = scale if scale is not None else 1.0 / math.sqrt(d)
    nkv = vt.shape[0]
    vhr = vt_head_rows if vt_head_rows > 0 else d
    Q = q.double().view(nimg, lq, heads, d).permute(0, 2, 1, 3)
    K = k.double().view(nkv, lkv, heads, d).permute(0, 2, 1, 3)
    V = vt.double().view(nkv, heads, vhr, vt.shape[2])[:, :, :d, :lkv].permute(0, 1, 3, 2)
    if vt_ones:
        assert bool((vt.view(nkv, heads, vhr, vt.shape[2])[:, :, d, :lkv] == 1).all()), "ones row missing"
    idx = torch.arange(nimg) // kv_div
    P = torch.softmax(Q @ K[idx].transpose(-1, -2) * scale, dim=-1)
    O = _h((P @ V[idx]).permute(0, 2, 1, 3).reshape(nimg * lq, heads * d))
    if out is not None:
        out.copy_(O)
        return out
    return O


def temporal_attention(qkv, *, nb, f_q, npix, heads, d, pe_q=None, kv=None, f_kv=None, f_kv_rank=0,
                       f_q_offset=0, kv_offsets=None, out=None):
    C = heads * d
    Q = qkv.double()[:, :C].view(nb, f_q, npix, heads, d)
    if pe_q is not None:
        Q = Q + pe_q[f_q_offset:f_q_offset + f_q].view(1, f_q, 1, heads, d)
    if kv is None:
        K = qkv.double()[:, C:2 * C].view(nb, f_q, npix, heads, d)
        V = qkv.double()[:, 2 * C:3 * C].view(nb, f_q, npix, heads, d)
    else:
        ko, vo = kv_offsets if kv_offsets is not None else (0, C)
        fr = f_kv_rank if f_kv_rank > 0 else f_kv
        G = f_kv // fr
        kvf = kv.double().view(G, nb, fr, npix, kv.shape[1])
        kvf = kvf.permute(1, 0, 2, 3, 4).reshape(nb, f_kv, npix, kv.shape[1])
        K = kvf[..., ko:ko + C].reshape(nb, f_kv, npix, heads, d)
        V = kvf[..., vo:vo + C].reshape(nb, f_kv, npix, heads, d)
    Qh, Kh, Vh = (t.permute(0, 2, 3, 1, 4) for t in (Q, K, V))           # [nb, npix, heads, f, d]
    P = torch.softmax(Qh @ Kh.transpose(-1, -2) / math.sqrt(d), dim=-1)
    O = _h((P @ Vh).permute(0, 3, 1, 2, 4).reshape(nb * f_q * npix, C))
    if out is not None:
        out.copy_(O)
        return out
    return O


def groupnorm(x0, gamma, beta, *, nimg, hw, groups, eps, silu, x1=None, out=None, ws=None, chunks=None):
    x = x0.double() if x1 is None else torch.cat([x0.double(), x1.double()], 1)
    C = x.shape[1]
    y = F.group_norm(x.view(nimg, hw, C).permute(0, 2, 1), groups, gamma.double(), beta.double(), eps)
    if silu:
        y = F.silu(y)
    y = _h(y.permute(0, 2, 1).reshape(nimg, hw, C))
    if chunks is None:
        return y.reshape(nimg * hw, C)
    G, pp = chunks                                   # exchange layout [G, nimg, pp, C], padding rows zero
    o = torch.zeros((nimg, G * pp, C), dtype=F16)
    o[:, :hw] = y
    return o.view(nimg, G, pp, C).permute(1, 0, 2, 3).reshape(G * nimg * pp, C).contiguous()


def unshard(back, *, nimg, hw, chunk_pix, x=None, out=None):
    C = back.shape[1]
    G = back.shape[0] // (nimg * chunk_pix)
    y = back.view(G, nimg, chunk_pix, C).permute(1, 0, 2, 3).reshape(nimg, G * chunk_pix, C)[:, :hw]
    y = y.reshape(nimg * hw, C)
    if x is not None:
        y = _h(y.double() + x.double())
    if out is None:
        return y.contiguous()
    out.copy_(y)
    return out


def layernorm(x, gamma, beta, *, eps=1e-5, add=None, add_row0=0, out=None, out2=None):
    y = F.layer_norm(x.double(), (x.shape[1],), gamma.double(), beta.double(), eps)
    if add is None:
        return _h(y)
    return _h(y), _h(y[add_row0:] + add.double())


def upsample2x(x, nimg, h, w):
    c = x.shape[1]
    y = x.view(nimg, h, w, c).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    return y.reshape(nimg * 4 * h * w, c).contiguous()


def im2col3x3(x, nimg, h, w, stride):
    c = x.shape[1]
    xi = x.double().view(nimg, h, w, c).permute(0, 3, 1, 2)
    cols = F.unfold(xi, 3, padding=1, stride=stride)                    # [nimg, c*9, L], row = ci*9 + tap
    L = cols.shape[2]
    cols = cols.view(nimg, c, 9, L).permute(0, 3, 2, 1).reshape(nimg * L, 9 * c)   # column tap*c + ci
    return _h(cols)


def time_embed(timestep, w1, b1, w2, b2, proj_w, proj_b, *, flip_sin_to_cos, freq_shift, scratch, out):
    dim = w1.shape[1]
    half = dim // 2
    expo = -math.log(10000.0) * torch.arange(half, dtype=torch.float64) / (half - freq_shift)
    e = timestep.reshape(-1)[:1].double()[:, None] * torch.exp(expo)[None]
    e = torch.cat([torch.sin(e), torch.cos(e)], -1)
    if flip_sin_to_cos:
        e = torch.cat([e[:, half:], e[:, :half]], -1)
    h = F.silu(e @ w1.double().t() + b1.double())
    h = F.silu(h @ w2.double().t() + b2.double())                          # SiLU applied by every resnet
    out.copy_((h @ proj_w.double().t())[0] + proj_b)
    return out


def latents_to_nhwc(sample, *, b, frame_idx, fl, cpad):
    b_src, c, Ft, h, w = sample.shape
    idx = torch.arange(fl) if frame_idx is None else frame_idx.long()
    x = sample[:, :, idx]                                                # [b_src, c, fl, h, w]
    x = x[torch.arange(b) % b_src].permute(0, 2, 3, 4, 1).reshape(b * fl * h * w, c)
    out = torch.zeros((b * fl * h * w, cpad), dtype=F16)
    out[:, :c] = x
    return out


def pred_accumulate(pred, acc, counter, *, frame_idx, fl):
    b, c, Ft, h, w = acc.shape
    idx = torch.arange(fl) if frame_idx is None else frame_idx.long()
    p = pred.double()[:, :c].view(b, fl, h, w, c).permute(0, 4, 1, 2, 3)      # [b, c, fl, h, w]
    for j in range(fl):                     # sequential: a window may list a frame once only, but keep += exact
        acc[:, :, idx[j]] += p[:, :, j]
        if counter is not None:
            counter[idx[j]] += 1.0


def cfg_ddim_step(acc, counter, latents, coef, guidance_scale, v_prediction):
    nb = acc.shape[0]
    eps = acc / counter.view(1, 1, -1, 1, 1)
    g = eps[0:1] + guidance_scale * (eps[1:2] - eps[0:1]) if nb == 2 else eps
    sa, sb, sap, sbp = [float(v) for v in coef]
    x = latents.double()
    if v_prediction:
        x0, e = sa * x - sb * g, sa * g + sb * x
    else:
        x0, e = (x - sb * g) / sa, g
    latents.copy_(_h(sap * x0 + sbp * e))


def cond_to_nhwc(x, *, c_first, c, ho, wo, cpad):
    nimg, ctot, h, w = x.shape
    iy = (torch.arange(ho) * h) // ho
    ix = (torch.arange(wo) * w) // wo
    s = x[:, c_first:c_first + c][:, :, iy][:, :, :, ix]                 # [nimg, c, ho, wo]
    out = torch.zeros((nimg * ho * wo, cpad), dtype=F16)
    out[:, :c] = s.permute(0, 2, 3, 1).reshape(nimg * ho * wo, c)
    return out


def softmax_rows_(x):
    x.copy_(_h(torch.softmax(x.double(), dim=-1)))
    return x


def im2col3x3_ex(x, nimg, h, w, stride, pad_lo):
    c = x.shape[1]
    xi = F.pad(x.double().view(nimg, h, w, c).permute(0, 3, 1, 2), (pad_lo, 1, pad_lo, 1))
    cols = F.unfold(xi, 3, padding=0, stride=stride)
    L = cols.shape[2]
    assert L == ((h + pad_lo - 2) // stride + 1) * ((w + pad_lo - 2) // stride + 1)
    return _h(cols.view(nimg, c, 9, L).permute(0, 3, 2, 1).reshape(nimg * L, 9 * c))


def quick_gelu_(x):
    xf = x.double()
    x.copy_(_h(xf * torch.sigmoid(1.702 * xf)))
    return x


def relu_(x):
    x.copy_(torch.relu(x))
    return x


def man_modulate(x, gb, *, nimg, hw, eps=1e-5):
    c = x.shape[1]
    xf = x.double().view(nimg, hw, c)
    mean = xf.mean(1, keepdim=True)
    var = xf.var(1, unbiased=False, keepdim=True)
    n = (xf - mean) * torch.rsqrt(var + eps)
    g = gb.double()[:, :c].view(nimg, hw, c)
    b = gb.double()[:, c:2 * c].view(nimg, hw, c)
    return _h((n * (1 + g) + b).reshape(nimg * hw, c))


_NAMES = ["gemm", "attention", "temporal_attention", "groupnorm", "layernorm", "upsample2x", "im2col3x3",
          "time_embed", "latents_to_nhwc", "cond_to_nhwc", "relu_", "man_modulate", "pred_accumulate",
          "cfg_ddim_step", "quick_gelu_", "softmax_rows_", "im2col3x3_ex", "unshard"]


def install(monkeypatch):
    """Route mikudance_b200.ops.* to the CPU contract above for the duration of one test."""
    from mikudance_b200 import ops
    g = globals()
    for n in _NAMES:
        monkeypatch.setattr(ops, n, g[n])


def engine_on_cpu(engine_cls, model):
    """An engine instance over CPU weights, skipping only the constructor's CUDA / fp16 device checks
    (the product constructor refuses CPU models)."""
    eng = engine_cls.__new__(engine_cls)
    eng._setup(model, torch.device("cpu"))
    return eng
