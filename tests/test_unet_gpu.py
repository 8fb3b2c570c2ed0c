"""GPU (-m gpu): the UNet3DConditionModel forward and the denoising loop on the sm_100a path against
the fp32 CPU oracle and against the golden vectors generated from the reference's own modules.

Tolerance for a whole UNet forward (about 300 fp16-stored ops deep): the north_star's "fp16 rtol 1e-3" is the
resolution of fp16 itself, which a 300-op-deep fp16 pipeline cannot hold against an fp32 result.  The yardstick
is measured, not modelled: the SAME oracle graph executed in real torch fp16 on the GPU (fp16 weights and
activations through ATen / cuDNN / SDPA — the arithmetic the reference itself runs with, scripts/
inference_video.py:66-69,95) against the fp32 oracle on the same inputs, in the same test.  The gate is
    rel_l2(ours, fp32 oracle) <= max(1.25 * rel_l2(torch-fp16 run of the oracle, fp32 oracle), 1e-3)   and <= 5e-3
i.e. we must be as close to the fp32 truth as the reference's own fp16 execution (measured values are printed)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _run_case(cfg, B, f, h, w, lctx, t, with_banks):
    from mikudance_b200 import synth
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from oracle import unet3d_oracle as O
    m, sd = D.build_model(cfg)
    x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=lctx)
    banks = synth.synthetic_banks(cfg, B * f, h, w) if with_banks else None
    ReferenceAttentionControl(m, mode="read", do_classifier_free_guidance=(B == 2), fusion_blocks="full")
    if with_banks:
        for blk, (name, c, ds) in zip(m.spatial_blocks(), synth.reader_bank_order(cfg)):
            blk.bank = [banks[name].to(D.DEV)]
    y = m(x.to(D.DEV, D.F16), torch.tensor(t), encoder_hidden_states=ctx.to(D.DEV, D.F16),
          return_dict=False)[0].float().cpu()
    sd32 = {k: v.float() for k, v in sd.items()}
    with torch.no_grad():
        yo = O.unet3d_forward(sd32, cfg, x.half().float(), t, ctx.half().float(), banks=banks,
                              cfg_guidance=(B == 2))
        # the reference's arithmetic: the oracle graph in torch fp16 on the GPU
        O.set_compute_dtype(torch.float16)
        try:
            sd16 = {k: v.to(D.DEV, D.F16) for k, v in sd.items()}
            b16 = {k: v.to(D.DEV, D.F16) for k, v in banks.items()} if banks is not None else None
            y16 = O.unet3d_forward(sd16, cfg, x.to(D.DEV, D.F16), t, ctx.to(D.DEV, D.F16), banks=b16,
                                   cfg_guidance=(B == 2)).float().cpu()
            del sd16, b16
        finally:
            O.set_compute_dtype(torch.float32)
    print(f"unet parity: ours vs fp32 oracle {_rel(y, yo):.3e}; torch-fp16 oracle vs fp32 oracle {_rel(y16, yo):.3e}")
    return y, yo, y16


@pytest.mark.parametrize("name", ["unet_tiny_cfg", "unet_tiny_nobank"])
def test_unet_tiny_vs_oracle_and_reference_golden(name):
    from mikudance_b200 import synth
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, f, h, w, lctx, t, with_banks = [int(v) for v in z["meta"]]
    y, yo, y16 = _run_case(synth.TINY_CONFIG, B, f, h, w, lctx, t, bool(with_banks))
    ref = torch.from_numpy(z["y"])                  # produced by the reference's own modules
    floor = _rel(y16, yo)
    assert _rel(yo, ref) < 1e-5
    assert _rel(y, yo) <= max(1.25 * floor, 1e-3) and _rel(y, yo) <= 5e-3, (_rel(y, yo), floor)
    assert _rel(y, ref) <= 5e-3
    assert (y - yo).abs().max().item() <= 5e-3 * yo.abs().max().item() + 5e-3


def test_unet_config_a_sd15_vs_oracle():
    """BASELINE config A: SD-1.5-sized UNet3D (1.31 B params), 32x32 latents, 4 frames, CFG, banks."""
    from mikudance_b200 import synth
    y, yo, y16 = _run_case(synth.SD15_CONFIG, 2, 4, 32, 32, 257, 499, True)
    floor = _rel(y16, yo)
    assert _rel(y, yo) <= max(1.25 * floor, 1e-3) and _rel(y, yo) <= 5e-3, (_rel(y, yo), floor)


def test_unet_config_b_shape_sd15_vs_oracle():
    """The benchmarked shape (BASELINE config B): SD-1.5-sized UNet3D, 96x96 latents (L0 self-attention over 9216
    tokens, d = 40: the two-stream / ones-row kernel, CTA-pair convolutions at M = 36 864 per image), 257 CLIP
    tokens, CFG and reference banks; 2 frames (4 images) so the fp32 CPU oracle finishes in well under a minute."""
    from mikudance_b200 import synth
    y, yo, y16 = _run_case(synth.SD15_CONFIG, 2, 2, 96, 96, 257, 499, True)
    floor = _rel(y16, yo)
    assert _rel(y, yo) <= max(1.25 * floor, 1e-3) and _rel(y, yo) <= 5e-3, (_rel(y, yo), floor)
    assert (y - yo).abs().max().item() <= 5e-3 * yo.abs().max().item() + 5e-3


def test_cfg_uncond_half_ignores_banks_on_gpu():
    """Domain property: scaling the banks changes the cond half only."""
    from mikudance_b200 import synth
    from mikudance_b200.reference_control import ReferenceAttentionControl
    cfg = synth.TINY_CONFIG
    m, _ = D.build_model(cfg)
    x, ctx = synth.synthetic_inputs(cfg, 2, 2, 8, 8, lctx=3)
    ReferenceAttentionControl(m, mode="read", do_classifier_free_guidance=True, fusion_blocks="full")
    outs = []
    for scale in (1.0, 3.0):
        banks = synth.synthetic_banks(cfg, 4, 8, 8, scale=scale)
        for blk, (name, c, ds) in zip(m.spatial_blocks(), synth.reader_bank_order(cfg)):
            blk.bank = [banks[name].to(D.DEV)]
        outs.append(m(x.to(D.DEV, D.F16), 10, ctx.to(D.DEV, D.F16), return_dict=False)[0])
    assert torch.equal(outs[0][0], outs[1][0])
    assert not torch.equal(outs[0][1], outs[1][1])


def test_denoise_loop_windows_graph_vs_oracle():
    """Two DDIM steps of a 6-frame clip with 4-frame windows (overlap 2 -> 3 windows, wrap-around),
    CUDA-graph replay vs eager (bit-identical) vs the oracle's restated pipeline loop."""
    from mikudance_b200 import synth
    from mikudance_b200.denoise import DenoiseLoop
    from mikudance_b200.scheduler import DDIMScheduler
    from oracle.ddim_oracle import DDIMOracle
    from oracle.pipeline_oracle import denoise_loop
    cfg = synth.TINY_CONFIG
    m, sd = D.build_model(cfg)
    F_, h, w, steps = 6, 8, 8, 2
    lat, ctx = synth.synthetic_inputs(cfg, 2, F_, h, w, lctx=5)
    lat = lat[:1].half()
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False, steps_offset=1,
              prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")

    def banks_for_window(wdw):
        return synth.synthetic_banks(cfg, 2 * len(wdw), h, w, seed=200 + wdw[0])

    results = []
    for use_graph in (True, False):
        loop = DenoiseLoop(m, DDIMScheduler(**kw), guidance_scale=3.5, context_frames=4, context_stride=1,
                           context_overlap=2, use_cuda_graph=use_graph)
        loop.prepare(lat.to(D.DEV).contiguous().clone(), ctx, steps, banks_for_window)
        assert [list(x) for x in loop.windows] == [[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 0, 1]]
        if use_graph:
            loop.capture()
        results.append(loop.run().float().cpu())
    assert torch.equal(results[0], results[1])
    sd32 = {k: v.float() for k, v in sd.items()}
    with torch.no_grad():
        want = denoise_loop(sd32, cfg, lat.float(), ctx.half().float(), steps, 3.5, banks_for_window,
                            context_frames=4, context_stride=1, context_overlap=2,
                            scheduler=DDIMOracle(**kw))
    assert _rel(results[0], want) <= 5e-3, _rel(results[0], want)


def test_packed_checkpoint_and_weight_cache_on_device(tmp_path, monkeypatch):
    """from_packed() (meta-device module + engine filled from one file) and a MDK_WEIGHT_CACHE
    hit run the same kernels on the same packed tensors: outputs are bit-identical to the freshly packed model."""
    from mikudance_b200 import synth
    from mikudance_b200.unet_3d import UNet3DConditionModel
    cfg = synth.TINY_CONFIG
    m, sd = D.build_model(cfg)
    x, ctx = synth.synthetic_inputs(cfg, 2, 3, 16, 16, lctx=9)
    x, ctx = x.to(D.DEV, D.F16), ctx.to(D.DEV, D.F16)
    y0 = m(x, torch.tensor(499), encoder_hidden_states=ctx, return_dict=False)[0]
    path = tmp_path / "unet.packed.safetensors"
    m.save_packed(path)
    p = UNet3DConditionModel.from_packed(path, device="cuda")
    assert all(q.device.type == "meta" for q in p.parameters()) and p.device.type == "cuda" and p.dtype == D.F16
    y1 = p(x, torch.tensor(499), encoder_hidden_states=ctx, return_dict=False)[0]
    assert torch.equal(y0, y1)
    monkeypatch.setenv("MDK_WEIGHT_CACHE", str(tmp_path / "cache"))
    m2, _ = D.build_model(cfg)
    assert m2.engine().weight_cache == "miss"
    m3, _ = D.build_model(cfg)
    assert m3.engine().weight_cache == "hit"
    y3 = m3(x, torch.tensor(499), encoder_hidden_states=ctx, return_dict=False)[0]
    assert torch.equal(y0, y3)
