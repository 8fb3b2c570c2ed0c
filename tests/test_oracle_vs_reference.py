"""CPU, only where /root/reference is mounted: the oracle restatement against the reference's own
modules imported unmodified through oracle/diffusers_standin (the pin for oracle/unet3d_oracle.py)."""
import os
import sys

import pytest
import torch

from conftest import REFERENCE, ROOT, have_reference

pytestmark = pytest.mark.skipif(not have_reference(), reason="/root/reference is not mounted here")


@pytest.fixture(scope="module")
def ref_env():
    saved = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_standin"))
    sys.path.insert(1, REFERENCE)
    sys.path.insert(2, os.path.join(ROOT, "oracle"))
    import make_golden
    yield make_golden
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.") or k.startswith("diffusers")]:
        del sys.modules[k]
    sys.modules.update(saved_mods)
    sys.path[:] = saved


@pytest.mark.parametrize("B,f,h,w,lctx,t,with_banks", [(2, 2, 8, 8, 5, 499, True), (1, 5, 8, 16, 3, 0, False)])
def test_oracle_equals_reference_modules(ref_env, B, f, h, w, lctx, t, with_banks):
    from mikudance_b200 import synth
    from oracle import unet3d_oracle as O
    cfg = synth.TINY_CONFIG
    model = ref_env.build_reference_unet(cfg)
    sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=3).items()}
    model.load_state_dict(sd)
    x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=lctx, seed=7)
    banks = synth.synthetic_banks(cfg, B * f, h, w, seed=9) if with_banks else None
    ref_env.install_banks(model, banks, cfg, do_cfg=(B == 2))
    with torch.no_grad():
        y_ref = model(x, torch.tensor(t), encoder_hidden_states=ctx, return_dict=False)[0]
        y = O.unet3d_forward(sd, cfg, x, t, ctx, banks=banks, cfg_guidance=(B == 2))
    assert ((y - y_ref).norm() / y_ref.norm()).item() < 1e-5
    # pairing order and state-dict contract
    assert set(model.state_dict()) == set(sd)


def test_uncond_half_ignores_bank(ref_env):
    """with CFG the uncond half never sees the bank (oracle property)."""
    from mikudance_b200 import synth
    from oracle import unet3d_oracle as O
    cfg = synth.TINY_CONFIG
    sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0).items()}
    x, ctx = synth.synthetic_inputs(cfg, 2, 2, 8, 8, lctx=3)
    b1 = synth.synthetic_banks(cfg, 4, 8, 8, seed=1)
    b2 = {k: 3.0 * v for k, v in b1.items()}
    with torch.no_grad():
        y1 = O.unet3d_forward(sd, cfg, x, 10, ctx, banks=b1)
        y2 = O.unet3d_forward(sd, cfg, x, 10, ctx, banks=b2)
    assert torch.equal(y1[0], y2[0])
    assert not torch.equal(y1[1], y2[1])
