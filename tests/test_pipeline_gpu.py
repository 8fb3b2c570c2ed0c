"""GPU (-m gpu): MikuDanceVideoPipeline / Pose2VideoPipeline call surface with stub VAE / CLIP /
reference-UNet modules (those stages are outside the hot path; the stubs only honour the interfaces
the reference pipeline calls).  The denoising loop inside is the real sm_100a path."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402


class _Dist:
    def __init__(self, mean):
        self.mean = mean


class _Enc:
    def __init__(self, mean):
        self.latent_dist = _Dist(mean)


class _Dec:
    def __init__(self, sample):
        self.sample = sample


class StubVAE(nn.Module):
    def __init__(self):
        super().__init__()
        self.config = type("C", (), {"block_out_channels": (1, 1, 1, 1)})()
        self.w = nn.Parameter(torch.randn(4, 3) * 0.5, requires_grad=False)

    def encode(self, x):
        z = torch.nn.functional.avg_pool2d(x.float(), 8)
        return _Enc(torch.einsum("oc,bchw->bohw", self.w.float(), z).to(x.dtype))

    def decode(self, z):
        y = torch.einsum("oc,bohw->bchw", self.w.float(), z.float())
        return _Dec(torch.nn.functional.interpolate(y, scale_factor=8.0).to(z.dtype))


class StubCLIP(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.vision_model = nn.Module()
        self.vision_model.post_layernorm = nn.LayerNorm(16)
        self.visual_projection = nn.Linear(16, dim, bias=False)
        self.patch = nn.Linear(3, 16)

    def forward(self, pixel_values):
        t = torch.nn.functional.avg_pool2d(pixel_values, 56).flatten(2).transpose(1, 2)   # [1, 16, 3]
        return type("O", (), {"last_hidden_state": self.patch(t)})()


class BasicTransformerBlock(nn.Module):          # name matters: ReferenceAttentionControl pairs by it
    def __init__(self, c, ds):
        super().__init__()
        self.norm1 = nn.LayerNorm(c)
        self.c, self.ds = c, ds
        self.bank = []


class StubReferenceUNet(nn.Module):
    """Writer: on forward, every block appends a deterministic bank [N, hw, C]."""

    def __init__(self, cfg):
        super().__init__()
        boc = cfg["block_out_channels"]
        self.down_blocks = nn.ModuleList([nn.ModuleList([BasicTransformerBlock(boc[i], 2 ** i) for _ in range(2)])
                                          for i in range(3)])
        self.up_blocks = nn.ModuleList([nn.ModuleList([BasicTransformerBlock(boc[3 - i], 2 ** (3 - i)) for _ in range(3)])
                                        for i in range(1, 4)])
        self.mid_block = nn.ModuleList([BasicTransformerBlock(boc[3], 8)])
        self.calls = 0

    def forward(self, x, t, encoder_hidden_states=None, return_dict=False):
        self.calls += 1
        n, _, h, w = x.shape
        for m in self.modules():
            if isinstance(m, BasicTransformerBlock):
                g = torch.Generator().manual_seed(m.c * 131 + m.ds)
                base = torch.randn((h // m.ds) * (w // m.ds), m.c, generator=g)
                m.bank.append((base[None] * x.float().mean(dim=(1, 2, 3)).cpu()[:, None, None].add(1.0)).to(x.device))
        return (None,)


def test_pipeline_call_surface_runs_and_hoists_reference_unet():
    from PIL import Image
    from mikudance_b200 import synth
    from mikudance_b200.scheduler import DDIMScheduler
    from src.pipelines.pipeline_mikudance import MikuDanceVideoPipeline
    from src.pipelines.pipeline_stage2_vdo import Pose2VideoPipeline
    cfg = synth.TINY_CONFIG
    unet, _ = D.build_model(cfg)
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False, steps_offset=1,
              prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")
    torch.manual_seed(0)
    vae, clip, ref = StubVAE().to(D.DEV).half(), StubCLIP(cfg["cross_attention_dim"]).to(D.DEV).half(), \
        StubReferenceUNet(cfg).to(D.DEV)
    F_, H, W = 5, 64, 64
    rng = np.random.RandomState(0)

    def img():
        return Image.fromarray(rng.randint(0, 255, (48, 40, 3), dtype=np.uint8))

    args = dict(ref_image=img(), ref_skel_image=img(), tgt_pose_images=[img() for _ in range(F_)],
                tgt_face_images=[img() for _ in range(F_)], tgt_hand_images=[img() for _ in range(F_)],
                scene_motion_npy=rng.randn(F_, 2, H // 8, W // 8).astype(np.float32), width=W, height=H,
                video_length=F_, num_inference_steps=3, guidance_scale=3.5)
    outs = []
    for cls, ctxf in ((MikuDanceVideoPipeline, 4), (Pose2VideoPipeline, 4)):
        pipe = cls(vae=vae, image_encoder=clip, reference_unet=ref, denoising_unet=unet,
                   scheduler=DDIMScheduler(**kw)).to(D.DEV)
        calls0 = ref.calls
        seen = []
        out = pipe(generator=torch.Generator().manual_seed(42), context_frames=ctxf, context_overlap=2,
                   callback=lambda i, t, lat: seen.append((i, int(t))), **args)
        vid = out.videos
        assert tuple(vid.shape) == (1, 3, F_, H, W) and vid.dtype == torch.float32
        assert torch.isfinite(vid).all() and float(vid.min()) >= 0.0 and float(vid.max()) <= 1.0
        n_windows = 3          # uniform(0, 3, 5, 4, 1, 2): [0..3], [2,3,4,0], [4,0,1,2]
        assert ref.calls - calls0 == n_windows          # once per window, not per step per window
        assert [t for _, t in seen] == [999, 666, 332]
        outs.append(vid)
    assert torch.equal(outs[0], outs[1])      # same defaults given explicitly -> identical results
