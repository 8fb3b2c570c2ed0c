"""CPU: the reference-UNet ("writer") oracle (oracle/refunet_oracle.py) against the
golden vectors generated from the reference's own `src/models/unet_2d_mix.py` (oracle/make_golden.py),
against that module itself where /root/reference is mounted, and the weight-container contract of
mikudance_b200.unet_2d_ref."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, ROOT, have_reference
from mikudance_b200 import synth
from oracle import refunet_oracle as R


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_refunet_oracle_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "refunet_tiny.npz"))
    N, h, w, lctx = [int(v) for v in z["meta"]]
    cfg = synth.TINY_CONFIG
    sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0, reference_unet=True).items()}
    x, ctx = synth.synthetic_reference_inputs(cfg, N, h, w, lctx=lctx)
    with torch.no_grad():
        y, banks = R.refunet_forward(sd, cfg, x.half().float(), 0, ctx.half().float())
    assert _rel(y, torch.from_numpy(z["y"])) < 1e-5
    order = [n for n, _, _ in synth.reader_bank_order(cfg)]
    assert len(order) == 16 and set(order) == set(banks)
    for i, n in enumerate(order):                      # all sixteen banks: norm and mean
        assert abs(float(banks[n].norm()) - z["bank_norm"][i]) <= 1e-5 * z["bank_norm"][i], n
        assert abs(float(banks[n].double().mean()) - z["bank_mean"][i]) <= 1e-5, n
    for key in z.files:                                # four banks element-wise
        if key.startswith("bank_") and key not in ("bank_norm", "bank_mean"):
            name = [n for n in order if "bank_" + n.replace(".", "_") == key][0]
            ref = torch.from_numpy(z[key])
            assert banks[name].shape == ref.shape
            assert _rel(banks[name], ref) < 1e-5 and (banks[name] - ref).abs().max().item() < 1e-4


def test_refunet_state_dict_contract():
    shapes = json.load(open(os.path.join(GOLDEN, "refunet_state_dict_sd15.json")))
    spec = {k: list(s) for k, s, _ in synth.state_dict_spec(synth.SD15_CONFIG, reference_unet=True)}
    assert len(shapes) == 706 and spec == shapes
    assert shapes["conv_in.weight"] == [320, 20, 3, 3]
    assert not any(k.startswith(("conv_out", "conv_norm_out")) or "motion_modules" in k for k in shapes)
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel
    cfg = synth.TINY_CONFIG
    m = UNet2DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"])
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: list(s) for k, s, _ in synth.state_dict_spec(cfg, reference_unet=True)}
    res = m.load_state_dict(synth.synthetic_state_dict(cfg, reference_unet=True), strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_from_unet_and_failure_modes():
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel, UNet2DWeights
    cfg = synth.TINY_CONFIG
    base_sd = {k: v for k, v in synth.synthetic_state_dict(cfg, seed=5, reference_unet=True).items()
               if not k.startswith("man_blocks")}
    base_sd["conv_in.weight"] = base_sd["conv_in.weight"][:, :4].contiguous()      # a plain 4-channel UNet
    base_sd["conv_out.weight"] = torch.zeros(4, 64, 3, 3, dtype=torch.float16)     # ignored by from_unet
    base = UNet2DWeights(dict(block_out_channels=cfg["block_out_channels"],
                              cross_attention_dim=cfg["cross_attention_dim"], in_channels=4), base_sd)
    new = UNet2DConditionModel.from_unet(base)                                     # unet_2d_mix.py:897-920
    sd = new.state_dict()
    assert torch.equal(sd["conv_in.weight"][:, :4], base_sd["conv_in.weight"].float())
    assert float(sd["conv_in.weight"][:, 4:].abs().max()) == 0.0
    for k in ("time_embedding.linear_2.bias", "down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.weight",
              "mid_block.resnets.1.conv2.weight", "up_blocks.3.resnets.2.conv_shortcut.weight"):
        assert torch.equal(sd[k], base_sd[k].float()), k
    with pytest.raises(RuntimeError, match="no CPU path"):
        new(torch.zeros(1, 22, 8, 8), 0, torch.zeros(1, 3, 64))
    with pytest.raises(RuntimeError, match="parameter container"):
        new.man_blocks[0].mlp_gamma(torch.zeros(1, 128, 8, 8))
    with pytest.raises(NotImplementedError):
        UNet2DConditionModel(use_linear_projection=True)
    with pytest.raises(RuntimeError):
        base(torch.zeros(1))


def test_man_modulation_only_flows_forward():
    """src/models/unet_2d_mix.py:1288-1289: the MAN block changes what flows on, the skip connections keep
    the un-modulated tensors — so zeroing every MAN conv leaves InstanceNorm(x) flowing and the motion map
    without influence."""
    cfg = synth.TINY_CONFIG
    sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0, reference_unet=True).items()}
    for k in list(sd):
        if k.startswith("man_blocks") and ("mlp_gamma" in k or "mlp_beta" in k):
            sd[k] = torch.zeros_like(sd[k])
    x, ctx = synth.synthetic_reference_inputs(cfg, 2, 16, 16, lctx=3)
    x2 = x.clone()
    x2[:, -2:] = 5.0 * x2[:, -2:] + 1.0
    with torch.no_grad():
        y1, b1 = R.refunet_forward(sd, cfg, x, 0, ctx)
        y2, b2 = R.refunet_forward(sd, cfg, x2, 0, ctx)
    assert torch.equal(y1, y2) and all(torch.equal(b1[k], b2[k]) for k in b1)


@pytest.mark.skipif(not have_reference(), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("N,h,w,lctx,one_ctx", [(3, 16, 24, 5, False), (2, 16, 16, 257, True)])
def test_refunet_oracle_equals_reference_module(N, h, w, lctx, one_ctx):
    saved = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_standin"))
    sys.path.insert(1, REFERENCE)
    sys.path.insert(2, os.path.join(ROOT, "oracle"))
    try:
        import make_golden
        cfg = synth.TINY_CONFIG
        model = make_golden.build_reference_refunet(cfg)
        sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=4, reference_unet=True).items()}
        assert set(model.state_dict()) == set(sd)
        model.load_state_dict(sd)
        x, ctx = synth.synthetic_reference_inputs(cfg, N, h, w, lctx=lctx, seed=11)
        if one_ctx:
            ctx = ctx[1:2]
        y_ref, b_ref = make_golden.run_reference_refunet(model, x, ctx.expand(N, -1, -1) if one_ctx else ctx, cfg)
        with torch.no_grad():
            y, banks = R.refunet_forward(sd, cfg, x, 0, ctx)
        assert _rel(y, y_ref) < 1e-5
        for k in b_ref:
            assert _rel(banks[k], b_ref[k]) < 1e-5, k
    finally:
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.") or k.startswith("diffusers")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved
