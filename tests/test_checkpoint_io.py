"""CPU: checkpoint / weight-format handling of the model containers — the on-disk layouts
scripts/inference_video.py reads: SD-1.5 `unet/config.json` + `diffusion_pytorch_model.{safetensors,bin}`, the motion
module `.ckpt` / `.safetensors`, `denoising_unet.pth` / `reference_unet.pth` state dicts, the VAE and CLIP folders."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from mikudance_b200 import synth


def _split_unet3d_sd(cfg):
    sd = synth.synthetic_state_dict(cfg, seed=3)
    motion = {k: v for k, v in sd.items() if "motion_modules" in k}
    spatial = {k: v for k, v in sd.items() if "motion_modules" not in k}
    return sd, spatial, motion


@pytest.mark.parametrize("fmt,mm_fmt", [("safetensors", "ckpt"), ("bin", "safetensors")])
def test_unet3d_from_pretrained_2d(tmp_path, fmt, mm_fmt):
    """src/models/unet_3d_mix.py:600-691: config.json -> 3-D block types, 2-D weights, motion checkpoint merged."""
    from mikudance_b200.unet_3d import UNet3DConditionModel
    cfg = synth.TINY_CONFIG
    sd, spatial, motion = _split_unet3d_sd(cfg)
    unet_dir = tmp_path / "sd15" / "unet"
    unet_dir.mkdir(parents=True)
    conf = dict(_class_name="UNet2DConditionModel", sample_size=64, in_channels=4, out_channels=4,
                block_out_channels=list(cfg["block_out_channels"]), layers_per_block=2, attention_head_dim=8,
                cross_attention_dim=cfg["cross_attention_dim"], norm_num_groups=32, norm_eps=1e-5, act_fn="silu",
                down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
                up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, mid_block_type="UNetMidBlock2DCrossAttn")
    (unet_dir / "config.json").write_text(json.dumps(conf))
    if fmt == "safetensors":
        save_file({k: v.contiguous() for k, v in spatial.items()}, str(unet_dir / "diffusion_pytorch_model.safetensors"))
    else:
        torch.save(spatial, str(unet_dir / "diffusion_pytorch_model.bin"))
    mm_path = tmp_path / f"mm.{mm_fmt}"
    if mm_fmt == "safetensors":
        save_file({k: v.contiguous() for k, v in motion.items()}, str(mm_path))
    else:
        torch.save(motion, str(mm_path))
    add = dict(use_inflated_groupnorm=True, use_motion_module=True, motion_module_resolutions=[1, 2, 4, 8],
               motion_module_mid_block=True, motion_module_type="Vanilla", unet_use_cross_frame_attention=False,
               unet_use_temporal_attention=False,
               motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                                         attention_block_types=["Temporal_Self", "Temporal_Self"],
                                         temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                         temporal_attention_dim_div=1))
    m = UNet3DConditionModel.from_pretrained_2d(str(tmp_path / "sd15"), str(mm_path), subfolder="unet",
                                                unet_additional_kwargs=add)
    got = m.state_dict()
    assert set(got) == set(sd)
    for k in ("conv_in.weight", "down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight",
              "up_blocks.2.motion_modules.1.temporal_transformer.proj_out.weight",
              "mid_block.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.1.to_q.weight"):
        assert torch.equal(got[k], sd[k].float()), k
    # mm_zero_proj_out=True drops the checkpoint's proj_out (kept zero-initialised, motion_module.py:73-76)
    m0 = UNet3DConditionModel.from_pretrained_2d(str(tmp_path / "sd15"), str(mm_path), subfolder="unet",
                                                 unet_additional_kwargs=add, mm_zero_proj_out=True)
    assert float(m0.state_dict()["mid_block.motion_modules.0.temporal_transformer.proj_out.weight"].abs().max()) == 0.0
    # the script then overlays denoising_unet.pth with strict=False (scripts/inference_video.py:111-114)
    pth = tmp_path / "denoising_unet.pth"
    torch.save({k: torch.zeros_like(v) for k, v in list(sd.items())[:5]}, str(pth))
    res = m.load_state_dict(torch.load(str(pth), map_location="cpu"), strict=False)
    assert not res.unexpected_keys and float(m.state_dict()["conv_in.weight"].abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="config.json"):
        UNet3DConditionModel.from_pretrained_2d(str(tmp_path / "nope"), str(mm_path), subfolder="unet")


def test_reference_unet_checkpoint_flow(tmp_path):
    """scripts/inference_video.py:81-85,115-117: base UNet folder -> from_unet -> reference_unet.pth (strict)."""
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel, UNet2DWeights
    cfg = synth.TINY_CONFIG
    rsd = synth.synthetic_state_dict(cfg, seed=4, reference_unet=True)
    base = {k: v for k, v in rsd.items() if not k.startswith("man_blocks")}
    base["conv_in.weight"] = base["conv_in.weight"][:, :4].contiguous()
    d = tmp_path / "sd15" / "unet"
    d.mkdir(parents=True)
    (d / "config.json").write_text(json.dumps(dict(block_out_channels=list(cfg["block_out_channels"]),
                                                   cross_attention_dim=cfg["cross_attention_dim"], in_channels=4)))
    save_file({k: v.contiguous() for k, v in base.items()}, str(d / "diffusion_pytorch_model.safetensors"))
    unet = UNet2DWeights.from_pretrained(str(tmp_path / "sd15"), subfolder="unet").to(device="cpu")
    ref = UNet2DConditionModel.from_unet(unet)
    torch.save(rsd, str(tmp_path / "reference_unet.pth"))
    res = ref.load_state_dict(torch.load(str(tmp_path / "reference_unet.pth"), map_location="cpu"))   # strict
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(ref.state_dict()["man_blocks.2.mlp_beta.weight"], rsd["man_blocks.2.mlp_beta.weight"].float())
    with pytest.raises(FileNotFoundError):
        (d / "diffusion_pytorch_model.safetensors").unlink()
        UNet2DWeights.from_pretrained(str(tmp_path / "sd15"), subfolder="unet")


def test_vae_and_clip_from_pretrained(tmp_path):
    from mikudance_b200.clip_vision import CLIPVisionModelWithProjection
    from mikudance_b200.vae import AutoencoderKL
    vcfg, ccfg = synth.TINY_VAE_CONFIG, synth.CLIP_TINY_CONFIG
    vdir, cdir = tmp_path / "vae", tmp_path / "image_encoder"
    vdir.mkdir()
    cdir.mkdir()
    vsd, csd = synth.synthetic_vae_state_dict(vcfg), synth.synthetic_clip_state_dict(ccfg)
    (vdir / "config.json").write_text(json.dumps(dict(_class_name="AutoencoderKL", **{k: (list(v) if isinstance(v, tuple) else v)
                                                                                    for k, v in vcfg.items()})))
    save_file({k: v.contiguous() for k, v in vsd.items()}, str(vdir / "diffusion_pytorch_model.safetensors"))
    vae = AutoencoderKL.from_pretrained(str(vdir))
    assert all(torch.equal(vae.state_dict()[k], v.float()) for k, v in vsd.items())
    (cdir / "config.json").write_text(json.dumps(dict(model_type="clip_vision_model", hidden_act="quick_gelu", **ccfg)))
    torch.save(csd, str(cdir / "pytorch_model.bin"))
    clip = CLIPVisionModelWithProjection.from_pretrained(str(cdir))
    assert all(torch.equal(clip.state_dict()[k], v.float()) for k, v in csd.items())
    tr = pytest.importorskip("transformers")
    hf = tr.CLIPVisionModelWithProjection(tr.CLIPVisionConfig(**ccfg))
    native = CLIPVisionModelWithProjection.from_transformers(hf)
    assert all(torch.equal(native.state_dict()[k], v) for k, v in hf.state_dict().items() if "position_ids" not in k)


def test_video_io_round_trip(tmp_path):
    """src.utils.util as imported by scripts/inference_video.py:24 (cv2-backed here; the reference uses PyAV)."""
    from src.utils.util import get_fps, read_frames, save_videos_grid
    vid = torch.zeros(2, 3, 5, 32, 48)
    for t in range(5):
        vid[0, :, t] = t / 4.0            # clip 0 brightens over time, clip 1 stays at 0.5
    vid[1] = 0.5
    path = str(tmp_path / "out" / "grid.mp4")
    save_videos_grid(vid, path, n_rows=2, fps=12)
    frames = read_frames(path)
    assert len(frames) == 5 and frames[0].size == (2 * 48 + 3 * 2, 32 + 2 * 2)
    assert float(get_fps(path)) == 12.0
    import numpy as np
    means = [np.asarray(f)[2:34, 2:50].mean() for f in frames]          # clip 0's cell
    assert means == sorted(means) and means[0] < 10 and means[-1] > 245
    assert abs(np.asarray(frames[2])[2:34, 52:100].mean() - 127.5) < 4   # clip 1's cell
    save_videos_grid(vid[:1], str(tmp_path / "one.gif"), fps=8)
    assert (tmp_path / "one.gif").stat().st_size > 0
    with pytest.raises(ValueError):
        save_videos_grid(vid, str(tmp_path / "x.avi"))
    with pytest.raises(FileNotFoundError):
        read_frames(str(tmp_path / "missing.mp4"))
