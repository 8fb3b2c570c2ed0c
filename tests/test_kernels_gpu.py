"""GPU (-m gpu): every sm_100a kernel against a plain PyTorch fp32 computation of the same op on the
same inputs, called through the C ABI (mikudance_b200.ops -> libmikudance_sm100.so).  Tolerance:
normalised L2 error <= 1e-3 per kernel for fp16 outputs (fp16 epsilon is 9.8e-4; observed 2.1e-4 to
2.9e-4 = one fp16 rounding of an fp32-accumulated result), exact for pure data movement."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402  (tests/ is on sys.path under pytest)


@pytest.fixture(autouse=True)
def _strict_tolerance(monkeypatch):
    orig = D.report

    def strict(name, got, ref, tol=1e-3):
        return orig(name, got, ref, tol=min(tol, 1e-3))
    monkeypatch.setattr(D, "report", strict)


def test_native_library_is_what_runs():
    from mikudance_b200 import _lib
    n0 = _lib.launch_count()
    a = torch.randn(256, 64, device="cuda").half()
    w = torch.randn(128, 64, device="cuda").half()
    D.ops.gemm(a, w)
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0 + 1


def test_gemm_shapes_and_tails():
    assert D.check_gemm_basic()


def test_gemm_fused_epilogues():
    assert D.check_gemm_epilogue()


def test_conv3x3_implicit_gemm():
    assert D.check_conv()


def test_groupnorm_layernorm():
    assert D.check_norms()


def test_groupnorm_bulk_copy_kernels(monkeypatch):
    """The cp.async.bulk-staged GroupNorm kernels (time-neutral, off by default) stay parity-green."""
    monkeypatch.setenv("MDK_GN_BULK", "1")
    assert D.check_norms()


def test_temporal_attention_incl_sharded_layout():
    assert D.check_temporal()


def test_glue_kernels_cfg_ddim_time_embed():
    assert D.check_misc()


def test_flash_attention_tcgen05():
    assert D.check_attn()


@pytest.mark.parametrize("env", [{"MDK_ATTN_2S": "0"}, {"MDK_ATTN_2S": "0", "MDK_ATTN_SK": "1"}, {"MDK_ATTN_PP": "3"}, {"MDK_ATTN_PP": "0"},
                                 {"MDK_ATTN_BKV": "64"}, {"MDK_ATTN_BKV": "128", "MDK_ATTN_PP": "0"},
                                 {"MDK_ATTN_POLY": "1"}, {"MDK_ATTN_STALE": "1"},
                                 {"MDK_ATTN_2S": "1"}, {"MDK_ATTN_2S": "1", "MDK_ATTN_POLY": "1"},
                                 {"MDK_ATTN_2S": "2"}, {"MDK_ATTN_2S": "2", "MDK_ATTN_POLY": "2"},
                                 {"MDK_ATTN_2S": "3"}, {"MDK_ATTN_2S": "3", "MDK_ATTN_POLY": "0"},
                                 {"MDK_ATTN_SPLITKV": "1"}, {"MDK_ATTN_SPLITKV": "1", "MDK_ATTN_POLY": "1"}])
def test_flash_attention_alternative_kernels(monkeypatch, env):
    """The kernels the dispatch heuristics do not pick by default (split-key, ping-pong for every
    head size, 64-key tiles) stay parity-green: the switches are read per call."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    assert D.check_attn()


@pytest.mark.parametrize("cg", ["1", "2"])
def test_gemm_single_cta_and_cta_pair_kernels(monkeypatch, cg):
    """Both GEMM kernels (128-row single CTA, 256-row cta_group::2 pair) on every shape / epilogue."""
    monkeypatch.setenv("MDK_GEMM_CG", cg)
    assert D.check_gemm_basic()
    assert D.check_gemm_epilogue()
    assert D.check_conv()


def test_gemm_b_stationary_kernel(monkeypatch):
    """The weight-stationary GEMM kernel (K <= 320, single CTA; picked by itself only for M >= 75 776) forced on every
    eligible shape / epilogue of the GEMM checks, plus the level-0 shape it is meant for."""
    monkeypatch.setenv("MDK_GEMM_BS", "2")
    monkeypatch.setenv("MDK_GEMM_CG", "1")
    assert D.check_gemm_basic()
    assert D.check_gemm_epilogue()
    assert D.check_gemm_bs_l0()


def test_argument_errors_are_reported():
    from mikudance_b200 import _lib
    a = torch.randn(64, 36, device="cuda").half()          # K = 36 is not a multiple of 8
    w = torch.randn(64, 36, device="cuda").half()
    with pytest.raises(_lib.MdkError, match="multiples of 8"):
        D.ops.gemm(a, w)
    with pytest.raises(TypeError):
        D.ops.gemm(a.float(), w)
