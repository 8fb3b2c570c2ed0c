"""CPU: host-side logic and the C-ABI boundary (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from mikudance_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mdk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mdk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mdk_abi_version() == 1
    assert lib.mdk_gemm_geglu_block() == 256


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    from mikudance_b200 import _lib, synth
    from mikudance_b200.unet_3d import UNet3DConditionModel
    lib = _lib.load_library()
    ctx = ctypes.c_void_p()
    assert lib.mdk_create(0, ctypes.byref(ctx)) != 0
    assert b"no CUDA device" in lib.mdk_last_error()
    with pytest.raises(_lib.MdkError):
        _lib.get_ctx(torch.device("cuda:0"))
    cfg = synth.TINY_CONFIG
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"], use_inflated_groupnorm=True,
                             use_motion_module=True, motion_module_mid_block=True,
                             motion_module_type="Vanilla", unet_use_cross_frame_attention=False,
                             unet_use_temporal_attention=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 4, 2, 8, 8), 1, torch.zeros(1, 3, 64))
    with pytest.raises(RuntimeError, match="parameter container"):
        m.conv_in(torch.zeros(1, 4, 8, 8))


def test_unsupported_configurations_raise():
    from mikudance_b200.unet_3d import UNet3DConditionModel
    with pytest.raises(NotImplementedError):
        UNet3DConditionModel(use_inflated_groupnorm=False, use_motion_module=True,
                             motion_module_type="Vanilla", motion_module_mid_block=True)
    with pytest.raises(NotImplementedError):
        UNet3DConditionModel(use_inflated_groupnorm=True, use_motion_module=False)


def test_reference_control_pairing_update_clear():
    from mikudance_b200 import synth
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from mikudance_b200.unet_3d import UNet3DConditionModel

    def build():
        cfg = synth.TINY_CONFIG
        return UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                                    cross_attention_dim=cfg["cross_attention_dim"],
                                    use_inflated_groupnorm=True, use_motion_module=True,
                                    motion_module_mid_block=True, motion_module_type="Vanilla",
                                    unet_use_cross_frame_attention=False,
                                    unet_use_temporal_attention=False)
    reader, writer = build(), build()
    rr = ReferenceAttentionControl(reader, mode="read", do_classifier_free_guidance=True,
                                   fusion_blocks="full")
    ww = ReferenceAttentionControl(writer, mode="write", fusion_blocks="full")
    names = {id(m): n for n, m in reader.named_modules()}
    order = [names[id(b)] for b in rr._blocks(reader)]
    assert order == [n + ".transformer_blocks.0" for n, _, _ in synth.reader_bank_order(synth.TINY_CONFIG)]
    assert order[5] == "mid_block.attentions.0.transformer_blocks.0"
    for i, b in enumerate(ww._blocks(writer)):
        b.bank.append(torch.full((2, 4, b.norm1.normalized_shape[0]), float(i)))
    rr.update(ww)
    for i, b in enumerate(rr._blocks(reader)):
        assert len(b.bank) == 1 and b.bank[0].dtype == torch.float16 and float(b.bank[0][0, 0, 0]) == i
    rr.clear()
    assert all(len(b.bank) == 0 for b in rr._blocks(reader))
    assert reader._ref_control["mode"] == "read" and reader._ref_control["do_classifier_free_guidance"]
    # midup selects mid + up blocks only
    r2 = ReferenceAttentionControl(build(), mode="read", fusion_blocks="midup")
    assert len(r2._blocks(r2.unet)) == 10


def test_reference_import_paths_resolve():
    from src.models.mutual_mix_attention import ReferenceAttentionControl  # noqa: F401
    from src.models.unet_3d_mix import UNet3DConditionModel  # noqa: F401
    from src.pipelines.context import get_context_scheduler
    assert list(get_context_scheduler("uniform")(0, 20, 16, 30, 1, 8)) == [list(range(16))]
    with pytest.raises(ValueError):
        get_context_scheduler("nope")


def test_geglu_panel_packing():
    from mikudance_b200.engine import _pack_geglu
    w = torch.arange(1024 * 2, dtype=torch.float32).reshape(1024, 2)
    b = torch.arange(1024, dtype=torch.float32)
    wp, bp = _pack_geglu(w, b, torch.device("cpu"), 256)
    assert bp[:128].tolist() == list(range(0, 128)) and bp[128:256].tolist() == list(range(512, 640))
    assert bp[256:384].tolist() == list(range(128, 256))
    assert torch.equal(wp[:, 0].float(), bp * 2)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """ABI drift guard: size and every field offset of the ctypes mirrors in mikudance_b200/_lib.py equal
    what a C compiler computes from include/mdk.h."""
    import shutil
    import subprocess
    from mikudance_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = [("mdk_gemm_args", _lib.GemmArgs), ("mdk_attn_args", _lib.AttnArgs), ("mdk_tattn_args", _lib.TattnArgs),
             ("mdk_gn_args", _lib.GnArgs), ("mdk_ln_args", _lib.LnArgs), ("mdk_temb_args", _lib.TembArgs),
             ("mdk_man_args", _lib.ManArgs)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mdk.h"', 'int main(void){']
    for cname, st in pairs:
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('return 0;}')
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mdk.h")).read(), flags=re.S)
    for cname, st in pairs:
        assert int(got[cname]) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"
        # and no field of the C struct is missing from the mirror
        body = [seg for seg in hdr.split("typedef struct {") if re.search(r"\}\s*" + cname + r"\s*;", seg)][0]
        body = body[:body.index("}")]
        c_fields = re.findall(r"([a-z_0-9]+)(?:\[\d+\])?\s*[;,]", body)
        assert set(c_fields) == {f for f, _ in st._fields_}, cname


def test_inference_script_import_surface(monkeypatch):
    """Every `src.*` name scripts/inference_video.py:17-24 imports resolves in this repository, and the opt-in
    `compat/diffusers` stand-in provides the three diffusers names of line 10."""
    import importlib
    import sys
    from src.models.unet_2d_condition import UNet2DConditionModel  # noqa: F401
    from src.models.unet_2d_mix import UNet2DConditionModel as MIX  # noqa: F401
    from src.models.unet_3d_mix import UNet3DConditionModel  # noqa: F401
    from src.pipelines.pipeline_mikudance import MikuDanceVideoPipeline  # noqa: F401
    from src.pipelines.pipeline_stage2_vdo import Pose2VideoPipeline  # noqa: F401
    from src.utils.util import get_fps, read_frames, save_videos_grid  # noqa: F401
    try:
        import diffusers  # noqa: F401
        pytest.skip("a real diffusers is installed: the stand-in must not be used")
    except ImportError:
        pass
    monkeypatch.syspath_prepend(os.path.join(ROOT, "compat"))
    try:
        d = importlib.import_module("diffusers")
        assert d.AutoencoderKL.__module__ == "mikudance_b200.vae"
        sch = d.DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False,
                              steps_offset=1, prediction_type="v_prediction", rescale_betas_zero_snr=True,
                              timestep_spacing="trailing")
        sch.set_timesteps(20)
        assert sch.timesteps.tolist()[:2] == [999, 949]
        with pytest.raises(NotImplementedError):
            d.AutoencoderKLTemporalDecoder.from_pretrained("x")
    finally:
        sys.modules.pop("diffusers", None)


def test_integration_md_binding_stub_matches_the_header_structs():
    """INTEGRATION.md's ctypes stub (what a reference maintainer would paste) declares mdk_gemm_args field for field as
    mikudance_b200/_lib.py does (which test_ctypes_structs_match_the_c_header ties to include/mdk.h), and only calls
    entry points the library exports."""
    import ctypes as C
    import re
    from mikudance_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(.*?)```", text, re.S).group(1)
    cls_src = re.search(r"(class GemmArgs\(C\.Structure\):.*?\n)\n", code, re.S).group(1)
    ns = {"C": C}
    exec(cls_src, ns)                                                      # the struct declaration only
    mine = [(n, t) for n, t in ns["GemmArgs"]._fields_]
    ref = [(n, t) for n, t in _lib.GemmArgs._fields_]
    assert [n for n, _ in mine] == [n for n, _ in ref]
    assert all(C.sizeof(a) == C.sizeof(b) for (_, a), (_, b) in zip(mine, ref))
    assert C.sizeof(ns["GemmArgs"]) == C.sizeof(_lib.GemmArgs)
    for sym in set(re.findall(r"lib\.(mdk_\w+)", code)):
        assert sym in _lib.EXPORTS, sym
