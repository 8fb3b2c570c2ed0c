"""Generate tests/golden/* from the REFERENCE's own modules (run in the build container, where
/root/reference is mounted; the GPU box has no /root/reference, so the fixtures are committed).

    python oracle/make_golden.py

What comes from where:
  * unet_tiny_*.npz      — outputs of the reference's unmodified `UNet3DConditionModel` +
                           `ReferenceAttentionControl(mode="read")` (src/models/unet_3d_mix.py,
                           src/models/mutual_mix_attention.py) imported from /root/reference through
                           oracle/diffusers_standin, fp32 CPU, on mikudance_b200.synth inputs/weights.
  * refunet_tiny.npz     — the reference's unmodified 2-D reference UNet (src/models/unet_2d_mix.py +
                           man_module.py) under ReferenceAttentionControl(mode="write"): its output
                           sample, four of the sixteen feature banks in full and the L2 norm / mean of all
                           sixteen (in the writer/reader pairing order); refunet_state_dict_sd15.json is
                           its key -> shape contract at the SD-1.5 size.
  * clip_tiny.npz        — outputs of transformers.CLIPVisionModelWithProjection (installed in this image) on
                           synthetic weights: last_hidden_state and the pipelines' image_prompt_embeds;
                           clip_state_dict_vitl14.json is its key -> shape contract at ViT-L/14 size.
  * context_windows.json — outputs of the reference's src/pipelines/context.py (imports untouched).
  * state_dict_sd15.json — key -> shape of the reference model built with the SD-1.5 config +
                           configs/inference/mikudance_config.yaml (the weight-container contract).
  * ddim_tables.json     — timestep tables / alphas from oracle.ddim_oracle (diffusers is absent:
                           restated, cross-checked against the values listed in SURVEY.md §8a16).
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "diffusers_standin"))
sys.path.insert(1, REF)
sys.path.insert(2, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def build_reference_unet(cfg):
    import yaml
    from src.models.unet_3d_mix import UNet3DConditionModel  # the reference's class
    kw = yaml.safe_load(open(os.path.join(REF, "configs/inference/mikudance_config.yaml")))
    kw = kw["unet_additional_kwargs"]
    base = dict(sample_size=64, in_channels=4, out_channels=4, center_input_sample=False,
                flip_sin_to_cos=True, freq_shift=0, block_out_channels=cfg["block_out_channels"],
                layers_per_block=2, downsample_padding=1, mid_block_scale_factor=1, act_fn="silu",
                norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=cfg["cross_attention_dim"],
                attention_head_dim=8)
    return UNet3DConditionModel(**base, **kw).eval()


def install_banks(model, banks, cfg, do_cfg):
    from src.models.attention import TemporalBasicTransformerBlock
    from src.models.mutual_mix_attention import ReferenceAttentionControl, torch_dfs
    from mikudance_b200 import synth
    ReferenceAttentionControl(model, mode="read", do_classifier_free_guidance=do_cfg,
                              fusion_blocks="full")
    mods = [m for m in torch_dfs(model) if isinstance(m, TemporalBasicTransformerBlock)]
    mods = sorted(mods, key=lambda x: -x.norm1.normalized_shape[0])
    for mod, (name, c, ds) in zip(mods, synth.reader_bank_order(cfg)):
        mod.bank = [banks[name].clone()] if banks is not None else []


REFUNET_FULL_BANKS = ("down_blocks.0.attentions.0", "up_blocks.1.attentions.2", "mid_block.attentions.0",
                      "up_blocks.3.attentions.2")


def build_reference_refunet(cfg):
    from src.models.unet_2d_mix import UNet2DConditionModel  # the reference's class
    return UNet2DConditionModel(block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"]).eval()


def run_reference_refunet(model, x, ctx, cfg):
    """-> (sample, {attention path: bank}) of the reference writer, banks keyed in pairing order."""
    from src.models.attention import BasicTransformerBlock
    from src.models.mutual_mix_attention import ReferenceAttentionControl, torch_dfs
    from mikudance_b200 import synth
    writer = ReferenceAttentionControl(model, mode="write", do_classifier_free_guidance=True,
                                       fusion_blocks="full", batch_size=1)
    with torch.no_grad():
        y = model(x, torch.zeros((), dtype=torch.int64), encoder_hidden_states=ctx, return_dict=False)[0]
    mods = [m for m in torch_dfs(model) if isinstance(m, BasicTransformerBlock)]
    mods = sorted(mods, key=lambda m: -m.norm1.normalized_shape[0])
    banks = {name: mod.bank[0].clone() for mod, (name, c, ds) in zip(mods, synth.reader_bank_order(cfg))}
    writer.clear()
    return y, banks


def main():
    from mikudance_b200 import synth
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    cfg = synth.TINY_CONFIG
    model = build_reference_unet(cfg)
    sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0).items()}
    model.load_state_dict(sd)
    cases = [("unet_tiny_cfg", 2, 4, 16, 16, 9, 949, True),
             ("unet_tiny_nobank", 1, 3, 8, 24, 5, 19, False)]
    for name, B, f, h, w, lctx, t, with_banks in cases:
        x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=lctx)
        x, ctx = x.half().float(), ctx.half().float()
        banks = synth.synthetic_banks(cfg, B * f, h, w) if with_banks else None
        install_banks(model, banks, cfg, do_cfg=(B == 2))
        with torch.no_grad():
            y = model(x, torch.tensor(t), encoder_hidden_states=ctx, return_dict=False)[0]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy().astype(np.float32),
                            meta=np.array([B, f, h, w, lctx, t, int(with_banks)]))
        print(name, tuple(y.shape), float(y.abs().mean()))

    # reference UNet (writer)
    ref = build_reference_refunet(cfg)
    rsd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0, reference_unet=True).items()}
    ref.load_state_dict(rsd)
    N, h, w, lctx = 2, 24, 24, 7
    x, ctx = synth.synthetic_reference_inputs(cfg, N, h, w, lctx=lctx)
    x, ctx = x.half().float(), ctx.half().float()
    y, banks = run_reference_refunet(ref, x, ctx, cfg)
    order = [n for n, _, _ in synth.reader_bank_order(cfg)]
    np.savez_compressed(
        os.path.join(OUT, "refunet_tiny.npz"), y=y.numpy().astype(np.float32),
        meta=np.array([N, h, w, lctx]),
        bank_norm=np.array([float(banks[n].norm()) for n in order], dtype=np.float64),
        bank_mean=np.array([float(banks[n].double().mean()) for n in order], dtype=np.float64),
        **{"bank_" + n.replace(".", "_"): banks[n].numpy().astype(np.float32) for n in REFUNET_FULL_BANKS})
    print("refunet_tiny", tuple(y.shape), float(y.abs().mean()), len(banks))
    bigref = build_reference_refunet(synth.SD15_CONFIG)
    rshapes = {k: list(v.shape) for k, v in bigref.state_dict().items()}
    json.dump(rshapes, open(os.path.join(OUT, "refunet_state_dict_sd15.json"), "w"), indent=0)
    print("reference-unet sd15 tensors", len(rshapes), "params", sum(int(np.prod(s)) for s in rshapes.values()))
    del bigref

    # CLIP image embedding (third-party transformers, installed here): outputs of the real model
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    ccfg = synth.CLIP_TINY_CONFIG
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(**ccfg)).eval()
    csd = {k: v.float() for k, v in synth.synthetic_clip_state_dict(ccfg, seed=0).items()}
    missing = clip.load_state_dict(csd, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    px = synth.synthetic_pixel_values(ccfg, 2).half().float()
    with torch.no_grad():
        o = clip(px)
        emb = clip.visual_projection(clip.vision_model.post_layernorm(o.last_hidden_state))   # pipeline :405-417
    np.savez_compressed(os.path.join(OUT, "clip_tiny.npz"), last_hidden_state=o.last_hidden_state.numpy(),
                        image_prompt_embeds=emb.numpy())
    big = CLIPVisionModelWithProjection(CLIPVisionConfig(**synth.CLIP_VITL14_CONFIG))
    json.dump({k: list(v.shape) for k, v in big.state_dict().items() if "position_ids" not in k},
              open(os.path.join(OUT, "clip_state_dict_vitl14.json"), "w"), indent=0)
    print("clip_tiny", tuple(emb.shape), float(emb.abs().mean()))
    del big

    spec = importlib.util.spec_from_file_location("refctx", os.path.join(REF, "src/pipelines/context.py"))
    refctx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refctx)
    rows = []
    for F_, cs, ov in [(4, 30, 8), (16, 30, 8), (32, 30, 8), (32, 32, 8), (64, 30, 8), (64, 32, 8),
                       (100, 30, 8), (65, 16, 4), (232, 30, 8)]:
        for step in (0, 1, 5):
            for stride in (1, 2, 3):
                rows.append(dict(step=step, num_frames=F_, context_size=cs, context_stride=stride,
                                 context_overlap=ov,
                                 windows=[list(map(int, w)) for w in
                                          refctx.uniform(step, 20, F_, cs, stride, ov)]))
    json.dump(rows, open(os.path.join(OUT, "context_windows.json"), "w"))
    print("context cases", len(rows))

    big = build_reference_unet(synth.SD15_CONFIG)
    shapes = {k: list(v.shape) for k, v in big.state_dict().items()}
    json.dump(shapes, open(os.path.join(OUT, "state_dict_sd15.json"), "w"), indent=0)
    print("sd15 tensors", len(shapes), "params", sum(int(np.prod(s)) for s in shapes.values()))

    from oracle.ddim_oracle import DDIMOracle
    sch = DDIMOracle()
    tab = dict(timesteps={str(n): [int(t) for t in sch.set_timesteps(n)] for n in (2, 20, 50)},
               alphas_cumprod={str(i): float(sch.alphas_cumprod[i]) for i in (0, 19, 49, 499, 949, 999)},
               source="oracle.ddim_oracle (restated diffusers 0.24.0 DDIMScheduler); SURVEY.md §8a16")
    json.dump(tab, open(os.path.join(OUT, "ddim_tables.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
