"""GPU (-m gpu): the native CLIP image encoder against the fp32 oracle (itself pinned to
transformers' implementation) and the golden vectors generated from transformers' model.

STATUS: validated on a B200 in round 2 (profiles/r02_first_call.log); collected by the default -m gpu run.  The host orchestration is covered on CPU by tests/test_clip_oracle.py; the kernels it
uses (GEMM with bias / row bias / transposed V^T segment, LayerNorm, attention at 257 tokens) are the validated
ones, plus the new in-place QuickGELU."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402
from conftest import GOLDEN  # noqa: E402


def test_clip_tiny_matches_oracle():
    assert D.check_clip()


def test_clip_matches_transformers_golden():
    from mikudance_b200 import synth
    from mikudance_b200.clip_vision import CLIPVisionModelWithProjection
    z = np.load(os.path.join(GOLDEN, "clip_tiny.npz"))
    cfg = synth.CLIP_TINY_CONFIG
    m = CLIPVisionModelWithProjection(**cfg)
    m.load_state_dict(synth.synthetic_clip_state_dict(cfg, seed=0))
    m = m.to(device=D.DEV, dtype=D.F16).eval()
    emb = m.image_prompt_embeds(synth.synthetic_pixel_values(cfg, 2).half().to(D.DEV)).float().cpu()
    ref = torch.from_numpy(z["image_prompt_embeds"])
    assert ((emb - ref).norm() / ref.norm()).item() < 5e-3


def test_clip_vit_l14_size():
    assert D.check_clip_vitl14()
