"""GPU (-m gpu): MikuDanceVideoPipeline with EVERY model stage native — CLIP image encoder, VAE encode of the
condition images, reference UNet (writer), denoising loop, VAE decode — against the same pipeline whose VAE / CLIP /
writer are fp32 PyTorch modules built from the oracles (same weights).

STATUS: validated on a B200 in round 2 (profiles/r02_first_call.log); collected by the default -m gpu run."""
import os

import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402


class OracleVAE(nn.Module):
    """diffusers-style VAE interface over oracle/vae_oracle.py (fp32, CPU)."""

    def __init__(self, cfg, sd):
        super().__init__()
        self.cfg, self.sd = cfg, {k: v.float() for k, v in sd.items()}
        self.config = type("C", (), {"block_out_channels": cfg["block_out_channels"]})()
        self.p = nn.Parameter(torch.zeros(1, dtype=torch.float16), requires_grad=False)

    def encode(self, x):
        from oracle import vae_oracle as V
        with torch.no_grad():
            m = V.encode_mean(self.sd, self.cfg, x.float().cpu()).to(x.device, x.dtype)
        return type("E", (), {"latent_dist": type("Dist", (), {"mean": m})()})()

    def decode(self, z):
        from oracle import vae_oracle as V
        with torch.no_grad():
            y = V.decode(self.sd, self.cfg, z.float().cpu()).to(z.device, z.dtype)
        return type("O", (), {"sample": y})()


class OracleCLIP(nn.Module):
    """transformers-style CLIP interface over oracle/clip_oracle.py (fp32, CPU)."""

    def __init__(self, cfg, sd):
        super().__init__()
        self.cfg, self.sd = cfg, {k: v.float() for k, v in sd.items()}
        self.p = nn.Parameter(torch.zeros(1, dtype=torch.float16), requires_grad=False)
        outer = self

        class _PLN(nn.Module):
            def forward(self, x):
                from oracle import clip_oracle as C
                return C._ln(outer.sd, "vision_model.post_layernorm", x.float().cpu(), cfg["layer_norm_eps"]).to(x.device, x.dtype)

        class _VP(nn.Module):
            def forward(self, x):
                return torch.nn.functional.linear(x.float().cpu(), outer.sd["visual_projection.weight"]).to(x.device, x.dtype)
        self.vision_model = nn.Module()
        self.vision_model.post_layernorm = _PLN()
        self.visual_projection = _VP()

    def forward(self, pixel_values):
        from oracle import clip_oracle as C
        with torch.no_grad():
            lh = C.clip_last_hidden_state(self.sd, self.cfg, pixel_values.float().cpu())
        return type("O", (), {"last_hidden_state": lh.to(pixel_values.device, pixel_values.dtype)})()


def test_fully_native_pipeline_matches_oracle_stages():
    import test_refunet_gpu as TR
    from PIL import Image
    from mikudance_b200 import _lib, synth
    from mikudance_b200.clip_vision import CLIPVisionModelWithProjection
    from mikudance_b200.scheduler import DDIMScheduler
    from mikudance_b200.vae import AutoencoderKL
    from src.pipelines.pipeline_mikudance import MikuDanceVideoPipeline
    cfg = synth.TINY_CONFIG
    vcfg = synth.TINY_VAE_CONFIG
    ccfg = dict(synth.CLIP_TINY_CONFIG, image_size=224, projection_dim=cfg["cross_attention_dim"])
    unet, _ = D.build_model(cfg)
    ref, rsd = D.build_refunet(cfg)
    vsd = synth.synthetic_vae_state_dict(vcfg)
    csd = synth.synthetic_clip_state_dict(ccfg)
    vae = AutoencoderKL(**vcfg)
    vae.load_state_dict(vsd)
    vae = vae.to(D.DEV, torch.float16).eval()
    clip = CLIPVisionModelWithProjection(**ccfg)
    clip.load_state_dict(csd)
    clip = clip.to(D.DEV, torch.float16).eval()
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False, steps_offset=1,
              prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")
    F_, H, W = 4, 128, 128
    rng = np.random.RandomState(0)

    def img():
        return Image.fromarray(rng.randint(0, 255, (48, 40, 3), dtype=np.uint8))

    args = dict(ref_image=img(), ref_skel_image=img(), tgt_pose_images=[img() for _ in range(F_)],
                tgt_face_images=[img() for _ in range(F_)], tgt_hand_images=[img() for _ in range(F_)],
                scene_motion_npy=rng.randn(F_, 2, H // 8, W // 8).astype(np.float32), width=W, height=H,
                video_length=F_, num_inference_steps=2, guidance_scale=3.5, context_frames=4, context_overlap=2)
    vids = []
    stages = [(vae, clip, ref),
              (OracleVAE(vcfg, vsd).to(D.DEV), OracleCLIP(ccfg, csd).to(D.DEV), TR._OracleWriter(cfg, rsd))]
    for v, c, r in stages:
        pipe = MikuDanceVideoPipeline(vae=v, image_encoder=c, reference_unet=r, denoising_unet=unet,
                                      scheduler=DDIMScheduler(**kw)).to(D.DEV)
        n0 = _lib.launch_count()
        vids.append(pipe(generator=torch.Generator().manual_seed(7), **args).videos)
        assert _lib.launch_count() > n0
    assert tuple(vids[0].shape) == (1, 3, F_, H, W) and torch.isfinite(vids[0]).all()
    rel = ((vids[0] - vids[1]).norm() / vids[1].norm()).item()
    assert rel < 3e-2, rel
