"""CPU: the CLIP image-embedding oracle (oracle/clip_oracle.py) against the real
`transformers.CLIPVisionModelWithProjection` (installed in this image) and the golden vectors generated from
it; the native container's state-dict contract; the ClipEngine host orchestration against the oracle through
the CPU statement of the kernel contracts."""
import json
import os

import numpy as np
import pytest
import torch

import ops_contract_cpu as K
from conftest import GOLDEN
from mikudance_b200 import synth
from oracle import clip_oracle as C


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_clip_oracle_matches_transformers_golden():
    z = np.load(os.path.join(GOLDEN, "clip_tiny.npz"))
    cfg = synth.CLIP_TINY_CONFIG
    sd = {k: v.float() for k, v in synth.synthetic_clip_state_dict(cfg, seed=0).items()}
    px = synth.synthetic_pixel_values(cfg, 2).half().float()
    with torch.no_grad():
        lh = C.clip_last_hidden_state(sd, cfg, px)
        emb = C.image_prompt_embeds(sd, cfg, px)
    assert _rel(lh, torch.from_numpy(z["last_hidden_state"])) < 1e-5
    assert _rel(emb, torch.from_numpy(z["image_prompt_embeds"])) < 1e-5


def test_clip_oracle_equals_installed_transformers():
    tr = pytest.importorskip("transformers")
    cfg = dict(synth.CLIP_TINY_CONFIG, num_hidden_layers=2, image_size=42)      # 3x3 patches + class token
    model = tr.CLIPVisionModelWithProjection(tr.CLIPVisionConfig(**cfg)).eval()
    sd = {k: v.float() for k, v in synth.synthetic_clip_state_dict(cfg, seed=5).items()}
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys)
    px = synth.synthetic_pixel_values(cfg, 3, seed=9)
    with torch.no_grad():
        o = model(px)
        ref = model.visual_projection(model.vision_model.post_layernorm(o.last_hidden_state))   # pipeline :405-417
        assert _rel(C.clip_last_hidden_state(sd, cfg, px), o.last_hidden_state) < 1e-6
        assert _rel(C.image_prompt_embeds(sd, cfg, px), ref) < 1e-6


def test_clip_state_dict_contract():
    from mikudance_b200.clip_vision import CLIPVisionModelWithProjection
    shapes = json.load(open(os.path.join(GOLDEN, "clip_state_dict_vitl14.json")))
    spec = {k: list(s) for k, s, _ in synth.clip_state_dict_spec(synth.CLIP_VITL14_CONFIG)}
    assert spec == shapes and len(shapes) == 392
    m = CLIPVisionModelWithProjection(**synth.CLIP_TINY_CONFIG)
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: list(s) for k, s, _ in synth.clip_state_dict_spec(synth.CLIP_TINY_CONFIG)}
    sd = synth.synthetic_clip_state_dict(synth.CLIP_TINY_CONFIG)
    sd["vision_model.embeddings.position_ids"] = torch.arange(17)[None]        # older checkpoints carry this buffer
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 56, 56))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.vision_model.post_layernorm(torch.zeros(1, 17, 128))
    with pytest.raises(NotImplementedError):
        CLIPVisionModelWithProjection(hidden_act="gelu")


def test_clip_engine_orchestration_matches_oracle(monkeypatch):
    from mikudance_b200.clip_vision import ClipEngine, CLIPVisionModelWithProjection
    K.install(monkeypatch)
    cfg = synth.CLIP_TINY_CONFIG
    sd = synth.synthetic_clip_state_dict(cfg, seed=0)
    m = CLIPVisionModelWithProjection(**cfg)
    m.load_state_dict(sd)
    m = m.half().eval()
    eng = ClipEngine.__new__(ClipEngine)
    eng._setup(m, torch.device("cpu"))          # the product constructor refuses CPU models
    px = synth.synthetic_pixel_values(cfg, 2).half()
    lh = eng.last_hidden_state(px)
    with torch.no_grad():
        want = C.clip_last_hidden_state({k: v.float() for k, v in sd.items()}, cfg, px.float())
    assert lh.shape == want.shape and _rel(lh, want) < 3e-3
    z = np.load(os.path.join(GOLDEN, "clip_tiny.npz"))
    assert _rel(lh, torch.from_numpy(z["last_hidden_state"])) < 3e-3
    with pytest.raises(ValueError, match="doesn't match model"):
        eng.last_hidden_state(torch.zeros(1, 3, 28, 28, dtype=torch.float16))
