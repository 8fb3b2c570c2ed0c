"""CPU: the VAE restatement (oracle/vae_oracle.py, SURVEY.md §8f row 2 — pinned to transformers' LDM encoder /
decoder implementations by tests/test_vae_oracle_pin.py, and to the SD-1.x VAE's parameter count here), the native
container's state-dict contract, and the
VaeEngine host orchestration against the restatement through the CPU statement of the kernel contracts."""
import pytest
import torch

import ops_contract_cpu as K
from mikudance_b200 import synth
from oracle import vae_oracle as V


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_vae_spec_matches_published_parameter_count():
    spec = synth.vae_state_dict_spec(synth.SD_VAE_CONFIG)
    assert len(spec) == 248
    assert sum(torch.Size(s).numel() for _, s, _ in spec) == 83_653_863      # AutoencoderKL of SD-1.x (sd-vae-ft-mse)
    from mikudance_b200.vae import AutoencoderKL
    m = AutoencoderKL(**{k: v for k, v in synth.TINY_VAE_CONFIG.items()})
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == \
        {k: list(s) for k, s, _ in synth.vae_state_dict_spec(synth.TINY_VAE_CONFIG)}
    assert 2 ** (len(m.config.block_out_channels) - 1) == 8                    # the pipelines' vae_scale_factor
    # checkpoints saved before diffusers 0.24 name the attention projections query/key/value/proj_attn
    sd = synth.synthetic_vae_state_dict(synth.TINY_VAE_CONFIG)
    old = {}
    for k, v in sd.items():
        for new, o in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            if f".attentions.0.{new}." in k:
                k = k.replace(f".attentions.0.{new}.", f".attentions.0.{o}.")
                v = v[:, :, None, None] if v.dim() == 2 else v
        old[k] = v
    res = m.load_state_dict(old, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["decoder.mid_block.attentions.0.to_v.weight"],
                       sd["decoder.mid_block.attentions.0.to_v.weight"].float())
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.encode(torch.zeros(1, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        AutoencoderKL(down_block_types=("DownEncoderBlock2D", "AttnDownEncoderBlock2D") * 2)


def test_vae_engine_orchestration_matches_restatement(monkeypatch):
    from mikudance_b200.vae import AutoencoderKL, DiagonalGaussianDistribution, VaeEngine
    K.install(monkeypatch)
    cfg = synth.TINY_VAE_CONFIG
    sd = synth.synthetic_vae_state_dict(cfg, seed=0)
    m = AutoencoderKL(**cfg)
    m.load_state_dict(sd)
    m = m.half().eval()
    eng = VaeEngine.__new__(VaeEngine)
    eng._setup(m, torch.device("cpu"))            # the product constructor refuses CPU models
    sd32 = {k: v.float() for k, v in sd.items()}
    x = synth._seeded_randn("vae_img", (2, 3, 64, 96), 1).half()              # non-square, mid block 8 x 12 tokens
    mom = eng.encode_moments(x)
    with torch.no_grad():
        want = V.encode_moments(sd32, cfg, x.float())
    assert mom.shape == want.shape == (2, 8, 8, 12)
    assert _rel(mom, want) < 5e-3
    dist = DiagonalGaussianDistribution(mom)
    assert torch.equal(dist.mean, mom[:, :4]) and dist.sample(torch.Generator().manual_seed(0)).shape == (2, 4, 8, 12)
    z = (0.5 * synth._seeded_randn("vae_lat", (2, 4, 8, 8), 2)).half()
    y = eng.decode(z)
    with torch.no_grad():
        yo = V.decode(sd32, cfg, z.float())
    assert y.shape == yo.shape == (2, 3, 64, 64)
    assert _rel(y, yo) < 5e-3
    with pytest.raises(ValueError):
        eng.encode_moments(torch.zeros(1, 3, 60, 64, dtype=torch.float16))
    with pytest.raises(NotImplementedError, match="% 8"):
        eng.decode(torch.zeros(1, 4, 2, 2, dtype=torch.float16))             # 4 mid-block tokens
