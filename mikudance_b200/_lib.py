"""ctypes binding of libmikudance_sm100.so (the C ABI declared in include/mdk.h).

The product path has no CPU fallback: if the shared library is missing, or there is no sm_100
device, every op raises.  Tensors are passed as raw device pointers; the caller keeps them alive
until the stream has consumed them (PyTorch's caching allocator + stream ordering does that for
tensors used on the current stream).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmikudance_sm100.so")

c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("a0", c_void_p), ("a1", c_void_p), ("b", c_void_p),
        ("lda0", c_int64), ("lda1", c_int64), ("ldb", c_int64),
        ("k0", c_int32), ("k1", c_int32),
        ("m", c_int32), ("n", c_int32),
        ("conv_taps", c_int32), ("nimg", c_int32), ("h", c_int32), ("w", c_int32),
        ("bias", c_void_p), ("row_bias", c_void_p),
        ("row_div", c_int32), ("row_mod", c_int32),
        ("residual", c_void_p), ("ldr", c_int64),
        ("geglu", c_int32), ("seg_cols", c_int32),
        ("out", c_void_p * 3), ("ldo", c_int64 * 3), ("out_trans", c_int32 * 3),
        ("trans_rows", c_int32), ("trans_ld", c_int64),
        ("trans_head_d", c_int32), ("trans_head_dp", c_int32),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("vt", c_void_p), ("out", c_void_p),
        ("ldq", c_int64), ("ldk", c_int64), ("ldvt", c_int64), ("ldo", c_int64),
        ("nimg", c_int32), ("nkv", c_int32), ("kv_div", c_int32),
        ("lq", c_int32), ("lkv", c_int32), ("heads", c_int32), ("d", c_int32),
        ("scale", c_float),
        ("vt_head_rows", c_int32), ("vt_ones", c_int32),
    ]


class TattnArgs(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("kv", c_void_p), ("pe_q", c_void_p), ("out", c_void_p),
        ("q_ld", c_int64), ("kv_ld", c_int64), ("out_ld", c_int64),
        ("q_off", c_int32), ("k_off", c_int32), ("v_off", c_int32),
        ("nb", c_int32), ("f_q", c_int32), ("f_kv", c_int32), ("f_kv_rank", c_int32),
        ("f_q_offset", c_int32), ("npix", c_int32), ("heads", c_int32), ("d", c_int32),
        ("scale", c_float),
    ]


class GnArgs(C.Structure):
    _fields_ = [
        ("x0", c_void_p), ("x1", c_void_p), ("c0", c_int32), ("c1", c_int32),
        ("nimg", c_int32), ("hw", c_int32), ("groups", c_int32), ("eps", c_float),
        ("gamma", c_void_p), ("beta", c_void_p), ("silu", c_int32),
        ("out", c_void_p), ("ws", c_void_p),
        ("out_chunk_pix", c_int32), ("out_chunks", c_int32),
    ]


class LnArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("rows", c_int64), ("c", c_int32), ("eps", c_float),
        ("gamma", c_void_p), ("beta", c_void_p), ("out", c_void_p),
        ("add", c_void_p), ("out2", c_void_p), ("add_row0", c_int64),
    ]


class TembArgs(C.Structure):
    _fields_ = [
        ("timestep", c_void_p), ("dim", c_int32), ("flip_sin_to_cos", c_int32),
        ("freq_shift", c_float),
        ("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p),
        ("edim", c_int32),
        ("proj_w", c_void_p), ("proj_b", c_void_p), ("nrows", c_int32),
        ("scratch", c_void_p), ("temb_out", c_void_p),
    ]


class ManArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("gb", c_void_p), ("ldgb", c_int64),
        ("nimg", c_int32), ("hw", c_int32), ("c", c_int32), ("eps", c_float),
        ("out", c_void_p), ("ws", c_void_p),
    ]


# every symbol include/mdk.h declares (tests check the library exports all of them)
EXPORTS = [
    "mdk_create", "mdk_destroy", "mdk_last_error", "mdk_abi_version", "mdk_launch_count",
    "mdk_gemm_f16", "mdk_gemm_geglu_block", "mdk_attn_fwd_f16", "mdk_temporal_attn_f16",
    "mdk_groupnorm_ws_bytes", "mdk_groupnorm_f16", "mdk_layernorm_f16", "mdk_upsample2x_f16",
    "mdk_im2col3x3_f16", "mdk_time_embed_f16", "mdk_latents_to_nhwc", "mdk_pred_accumulate",
    "mdk_cfg_ddim_step",
    "mdk_cond_to_nhwc_f16", "mdk_relu_f16", "mdk_man_ws_bytes", "mdk_man_modulate_f16",
    "mdk_attn_debug_trace", "mdk_quick_gelu_f16", "mdk_softmax_rows_f16", "mdk_im2col3x3_ex_f16",
    "mdk_unshard_add_f16",
]

_lib: Optional[C.CDLL] = None


class MdkError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MdkError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
            "mikudance_b200 has no CPU/PyTorch fallback path")
    lib = C.CDLL(LIB_PATH)
    lib.mdk_last_error.restype = C.c_char_p
    lib.mdk_launch_count.restype = c_int64
    lib.mdk_groupnorm_ws_bytes.restype = c_int64
    lib.mdk_groupnorm_ws_bytes.argtypes = [c_int32, c_int32]
    lib.mdk_create.argtypes = [C.c_int, C.POINTER(c_void_p)]
    lib.mdk_destroy.argtypes = [c_void_p]
    for name, argt in [("mdk_gemm_f16", GemmArgs), ("mdk_attn_fwd_f16", AttnArgs),
                       ("mdk_temporal_attn_f16", TattnArgs), ("mdk_groupnorm_f16", GnArgs),
                       ("mdk_layernorm_f16", LnArgs), ("mdk_time_embed_f16", TembArgs)]:
        fn = getattr(lib, name)
        fn.argtypes = [c_void_p, C.POINTER(argt), c_void_p]
        fn.restype = C.c_int
    lib.mdk_upsample2x_f16.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                       c_int32, c_void_p]
    lib.mdk_im2col3x3_f16.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                      c_int32, c_int32, c_int32, c_void_p]
    lib.mdk_latents_to_nhwc.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                        c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p]
    lib.mdk_pred_accumulate.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                        c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p]
    lib.mdk_unshard_add_f16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                        c_int32, c_void_p]
    lib.mdk_cfg_ddim_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                      c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
    lib.mdk_cond_to_nhwc_f16.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int32] * 9 + [c_void_p]
    lib.mdk_attn_debug_trace.argtypes = [c_void_p, c_int32]
    lib.mdk_quick_gelu_f16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.mdk_softmax_rows_f16.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p]
    lib.mdk_im2col3x3_ex_f16.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int32] * 7 + [c_void_p]
    lib.mdk_relu_f16.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.mdk_man_ws_bytes.restype = c_int64
    lib.mdk_man_ws_bytes.argtypes = [c_int32, c_int32]
    lib.mdk_man_modulate_f16.argtypes = [c_void_p, C.POINTER(ManArgs), c_void_p]
    lib.mdk_man_modulate_f16.restype = C.c_int
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().mdk_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise MdkError(f"{what}: {last_error()}")


_ctx_cache: dict = {}


def get_ctx(device: torch.device) -> c_void_p:
    """Per-device mdk_ctx (raises when there is no sm_100 GPU: no fallback)."""
    lib = load_library()
    if not torch.cuda.is_available():
        raise MdkError("mikudance_b200 needs a CUDA device (sm_100a); none is visible")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ctx = _ctx_cache.get(idx)
    if ctx is None:
        ctx = c_void_p()
        check(lib.mdk_create(idx, C.byref(ctx)), "mdk_create")
        _ctx_cache[idx] = ctx
    return ctx


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def cur_stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(load_library().mdk_launch_count())
