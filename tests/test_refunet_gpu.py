"""GPU (-m gpu): the native reference UNet ("writer") — its three own kernels
against plain PyTorch, the whole write-mode forward against the fp32 oracle and the golden vectors
generated from the reference's module, and the writer -> reader hand-over into the denoising UNet.

Validated on a B200 (profiles/r01_refunet_gpu.log): all tests below pass, including the SD-1.5-size case."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_diag as D  # noqa: E402
from conftest import GOLDEN  # noqa: E402


def test_refunet_own_kernels():
    assert D.check_refunet_ops()


def test_refunet_write_mode_matches_oracle():
    assert D.check_refunet_tiny()


def test_refunet_matches_reference_golden():
    from mikudance_b200 import synth
    z = np.load(os.path.join(GOLDEN, "refunet_tiny.npz"))
    N, h, w, lctx = [int(v) for v in z["meta"]]
    cfg = synth.TINY_CONFIG
    m, _ = D.build_refunet(cfg, seed=0)
    from mikudance_b200.reference_control import ReferenceAttentionControl
    writer = ReferenceAttentionControl(m, mode="write", do_classifier_free_guidance=True, fusion_blocks="full")
    x, ctx = synth.synthetic_reference_inputs(cfg, N, h, w, lctx=lctx)
    y = m(x.to(D.DEV, D.F16), 0, encoder_hidden_states=ctx.to(D.DEV, D.F16), return_dict=False)[0]
    torch.cuda.synchronize()
    ref = torch.from_numpy(z["y"])
    assert ((y.float().cpu() - ref).norm() / ref.norm()).item() < 5e-3
    order = [n for n, _, _ in synth.reader_bank_order(cfg)]
    for i, (blk, name) in enumerate(zip(writer._blocks(m), order)):
        b = blk.bank[0].float().cpu()
        assert abs(float(b.norm()) - z["bank_norm"][i]) <= 5e-3 * z["bank_norm"][i], name
        key = "bank_" + name.replace(".", "_")
        if key in z.files:
            g = torch.from_numpy(z[key])
            assert ((b - g).norm() / g.norm()).item() < 5e-3, name


def test_writer_to_reader_handover_equals_oracle_chain():
    """reference UNet (write) -> ReferenceAttentionControl.update -> denoising UNet (read): the product
    chain against the oracle chain refunet_oracle -> fp16 banks -> unet3d_oracle."""
    from mikudance_b200 import synth
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from oracle import refunet_oracle as R
    from oracle import unet3d_oracle as O
    cfg = synth.TINY_CONFIG
    f, h, w, lctx, t = 3, 32, 32, 9, 499
    ref, rsd = D.build_refunet(cfg, seed=0)
    den, dsd = D.build_model(cfg, seed=0)
    writer = ReferenceAttentionControl(ref, mode="write", do_classifier_free_guidance=True, fusion_blocks="full")
    reader = ReferenceAttentionControl(den, mode="read", do_classifier_free_guidance=True, fusion_blocks="full")
    cond, rctx = synth.synthetic_reference_inputs(cfg, 2 * f, h, w, lctx=lctx)
    x, ctx = synth.synthetic_inputs(cfg, 2, f, h, w, lctx=lctx)
    ref(cond.to(D.DEV, D.F16), 0, encoder_hidden_states=rctx.to(D.DEV, D.F16), return_dict=False)
    reader.update(writer)
    y = den(x.to(D.DEV, D.F16), torch.tensor(t), encoder_hidden_states=ctx.to(D.DEV, D.F16),
            return_dict=False)[0]
    torch.cuda.synchronize()
    with torch.no_grad():
        _, banks = R.refunet_forward({k: v.float() for k, v in rsd.items()}, cfg, cond.half().float(), 0,
                                     rctx.half().float())
        banks = {k: v.half() for k, v in banks.items()}                 # update() casts to fp16 (:353)
        yo = O.unet3d_forward({k: v.float() for k, v in dsd.items()}, cfg, x.half().float(), t,
                              ctx.half().float(), banks=banks, cfg_guidance=True)
    rel = ((y.float().cpu() - yo).norm() / yo.norm()).item()
    assert rel < 5e-3, rel


def test_refunet_sd15_size_config_a():
    assert D.check_refunet_a()


class _OracleWriter(torch.nn.Module):
    """Test double with the reference UNet's interface whose banks come from the fp32 ORACLE on the CPU
    (same block names / DFS order, so ReferenceAttentionControl pairs it like the real writer)."""

    def __init__(self, cfg, sd):
        super().__init__()
        import test_pipeline_gpu as TP
        self.cfg, self.sd = cfg, {k: v.float() for k, v in sd.items()}
        boc = cfg["block_out_channels"]
        B = TP.BasicTransformerBlock
        self.down_blocks = torch.nn.ModuleList([torch.nn.ModuleList([B(boc[i], 2 ** i) for _ in range(2)])
                                                for i in range(3)])
        self.up_blocks = torch.nn.ModuleList([torch.nn.ModuleList([B(boc[3 - i], 2 ** (3 - i)) for _ in range(3)])
                                              for i in range(1, 4)])
        self.mid_block = torch.nn.ModuleList([B(boc[3], 8)])
        self.calls = 0

    def forward(self, x, t, encoder_hidden_states=None, return_dict=False):
        from mikudance_b200 import synth
        from mikudance_b200.reference_control import torch_dfs
        from oracle import refunet_oracle as R
        self.calls += 1
        with torch.no_grad():
            _, banks = R.refunet_forward(self.sd, self.cfg, x.float().cpu(), 0, encoder_hidden_states.float().cpu())
        mods = [m for m in torch_dfs(self) if type(m).__name__ == "BasicTransformerBlock"]
        mods = sorted(mods, key=lambda m: -m.norm1.normalized_shape[0])
        for m, (name, c, ds) in zip(mods, synth.reader_bank_order(self.cfg)):
            m.bank.append(banks[name].to(x.device))
        return (None,)


def test_pipeline_with_native_reference_unet_matches_oracle_writer():
    """MikuDanceVideoPipeline end to end (stub VAE / CLIP) with the NATIVE reference UNet against the same
    pipeline whose writer is the CPU oracle: same videos within the fp16 noise of the bank values."""
    import test_pipeline_gpu as TP
    from PIL import Image
    from mikudance_b200 import _lib, synth
    from mikudance_b200.scheduler import DDIMScheduler
    from src.pipelines.pipeline_mikudance import MikuDanceVideoPipeline
    cfg = synth.TINY_CONFIG
    unet, _ = D.build_model(cfg)
    ref, rsd = D.build_refunet(cfg)
    kw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", clip_sample=False, steps_offset=1,
              prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing")
    torch.manual_seed(0)
    vae, clip = TP.StubVAE().to(D.DEV).half(), TP.StubCLIP(cfg["cross_attention_dim"]).to(D.DEV).half()
    F_, H, W = 5, 128, 128
    rng = np.random.RandomState(0)

    def img():
        return Image.fromarray(rng.randint(0, 255, (48, 40, 3), dtype=np.uint8))

    args = dict(ref_image=img(), ref_skel_image=img(), tgt_pose_images=[img() for _ in range(F_)],
                tgt_face_images=[img() for _ in range(F_)], tgt_hand_images=[img() for _ in range(F_)],
                scene_motion_npy=rng.randn(F_, 2, H // 8, W // 8).astype(np.float32), width=W, height=H,
                video_length=F_, num_inference_steps=3, guidance_scale=3.5, context_frames=4, context_overlap=2)
    vids = []
    for writer in (ref, _OracleWriter(cfg, rsd)):
        pipe = MikuDanceVideoPipeline(vae=vae, image_encoder=clip, reference_unet=writer, denoising_unet=unet,
                                      scheduler=DDIMScheduler(**kw)).to(D.DEV)
        n0 = _lib.launch_count()
        vids.append(pipe(generator=torch.Generator().manual_seed(42), **args).videos)
        assert _lib.launch_count() > n0
    assert tuple(vids[0].shape) == (1, 3, F_, H, W) and torch.isfinite(vids[0]).all()
    rel = ((vids[0] - vids[1]).norm() / vids[1].norm()).item()
    assert rel < 2e-2, rel
