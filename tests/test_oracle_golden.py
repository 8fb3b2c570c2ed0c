"""CPU: the oracle (oracle/) against the committed golden vectors generated from the reference's own
modules (oracle/make_golden.py), and the contract fixtures (state-dict keys, window lists, DDIM)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from mikudance_b200 import synth
from oracle import unet3d_oracle as O
from oracle.context_oracle import uniform as oracle_uniform
from oracle.ddim_oracle import DDIMOracle


@pytest.mark.parametrize("name", ["unet_tiny_cfg", "unet_tiny_nobank"])
def test_oracle_unet_matches_reference_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    B, f, h, w, lctx, t, with_banks = [int(v) for v in z["meta"]]
    cfg = synth.TINY_CONFIG
    sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0).items()}
    x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=lctx)
    banks = synth.synthetic_banks(cfg, B * f, h, w) if with_banks else None
    with torch.no_grad():
        y = O.unet3d_forward(sd, cfg, x.half().float(), t, ctx.half().float(), banks=banks,
                             cfg_guidance=(B == 2))
    ref = torch.from_numpy(z["y"])
    rel = ((y - ref).norm() / ref.norm()).item()
    assert rel < 1e-5, rel                       # fp32 vs fp32, different op order only
    assert (y - ref).abs().max().item() < 1e-4


def test_context_windows_match_reference_golden():
    rows = json.load(open(os.path.join(GOLDEN, "context_windows.json")))
    from mikudance_b200.context import uniform
    for r in rows:
        args = (r["step"], 20, r["num_frames"], r["context_size"], r["context_stride"], r["context_overlap"])
        assert oracle_uniform(*args) == r["windows"]          # oracle, bit exact
        assert list(uniform(*args)) == r["windows"]           # product host code, bit exact


def test_ddim_tables():
    tab = json.load(open(os.path.join(GOLDEN, "ddim_tables.json")))
    from mikudance_b200.scheduler import DDIMScheduler
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False,
              steps_offset=1, prediction_type="v_prediction", rescale_betas_zero_snr=True,
              timestep_spacing="trailing")
    orc = DDIMOracle(**kw)
    sch = DDIMScheduler(**kw)
    for n, ts in tab["timesteps"].items():
        assert [int(t) for t in orc.set_timesteps(int(n))] == ts          # index tables: ==
        sch.set_timesteps(int(n))
        assert sch.timesteps.tolist() == ts
    # values quoted in SURVEY.md §8a16
    survey = {999: 0.0, 949: 2.014479e-4, 499: 1.425294e-1, 49: 9.431506e-1, 19: 9.803064e-1, 0: 9.9915e-1}
    for i, v in survey.items():
        assert abs(float(orc.alphas_cumprod[i]) - v) <= 1e-6 + 1e-6 * v
        assert float(sch.alphas_cumprod[i]) == float(orc.alphas_cumprod[i])
    for i, v in tab["alphas_cumprod"].items():
        assert float(orc.alphas_cumprod[int(i)]) == v
    sch.set_timesteps(20)
    orc.set_timesteps(20)
    for t in (999, 949, 49):
        coef, prev = sch.step_coefficients(t)
        p2, a_t, a_prev = orc.coefficients(t)
        assert prev == p2 == t - 50
        assert torch.allclose(coef, torch.stack([a_t.sqrt(), (1 - a_t).sqrt(), a_prev.sqrt(),
                                                 (1 - a_prev).sqrt()]).float())


def test_state_dict_contract_sd15():
    """Key set and shapes of the reference UNet3DConditionModel (SD-1.5 + motion module)."""
    shapes = json.load(open(os.path.join(GOLDEN, "state_dict_sd15.json")))
    spec = {k: list(s) for k, s, _ in synth.state_dict_spec(synth.SD15_CONFIG)}
    assert len(shapes) == 1274
    assert spec == shapes


def test_model_state_dict_keys_tiny():
    from mikudance_b200.unet_3d import UNet3DConditionModel
    cfg = synth.TINY_CONFIG
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"], use_inflated_groupnorm=True,
                             use_motion_module=True, motion_module_mid_block=True,
                             motion_module_type="Vanilla",
                             motion_module_kwargs=dict(temporal_position_encoding=True,
                                                       temporal_position_encoding_max_len=32),
                             unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    want = {k: list(s) for k, s, _ in synth.state_dict_spec(cfg)}
    assert got == want
    # zero-initialised motion proj_out, like the reference constructor (motion_module.py:73-76)
    w = m.state_dict()["mid_block.motion_modules.0.temporal_transformer.proj_out.weight"]
    assert float(w.abs().max()) == 0.0
    res = m.load_state_dict(synth.synthetic_state_dict(cfg), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
