"""CPU: the product's HOST orchestration (UNetEngine / RefUNetEngine: weight packing, layouts, fused
q|k|v and GEGLU panels, skip concat, bank add / bank capture, MAN wiring, PE row bias) against the
oracle, with every kernel call routed to the CPU statement of its C-ABI contract
(tests/ops_contract_cpu.py).  The kernels themselves are covered by the `-m gpu` tests."""
import pytest
import torch

import ops_contract_cpu as K


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _unet3d(cfg):
    from mikudance_b200.unet_3d import UNet3DConditionModel
    return UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                                cross_attention_dim=cfg["cross_attention_dim"], use_inflated_groupnorm=True,
                                use_motion_module=True, motion_module_mid_block=True,
                                motion_module_type="Vanilla", unet_use_cross_frame_attention=False,
                                unet_use_temporal_attention=False)


@pytest.mark.parametrize("B,f,h,w,lctx,with_banks", [(2, 3, 16, 8, 9, True), (1, 2, 8, 8, 5, False),
                                                      (1, 1, 8, 8, 3, True),        # a single frame
                                                      (1, 32, 8, 8, 3, False)])     # the PE table's full length
def test_unet3d_engine_orchestration_matches_oracle(monkeypatch, B, f, h, w, lctx, with_banks):
    from mikudance_b200 import synth
    from mikudance_b200.engine import UNetEngine
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from oracle import unet3d_oracle as O
    K.install(monkeypatch)
    cfg = synth.TINY_CONFIG
    sd = synth.synthetic_state_dict(cfg, seed=1)
    model = _unet3d(cfg)
    model.load_state_dict(sd)
    model = model.half().eval()
    x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=lctx)
    banks = synth.synthetic_banks(cfg, B * f, h, w) if with_banks else None
    ReferenceAttentionControl(model, mode="read", do_classifier_free_guidance=(B == 2), fusion_blocks="full")
    if banks is not None:
        for blk, (name, c, ds) in zip(model.spatial_blocks(), synth.reader_bank_order(cfg)):
            blk.bank = [banks[name]]
    eng = K.engine_on_cpu(UNetEngine, model)
    y = eng.forward_api(x.half(), 499, ctx.half())
    with torch.no_grad():
        yo = O.unet3d_forward({k: v.float() for k, v in sd.items()}, cfg, x.half().float(), 499,
                              ctx.half().float(), banks=banks, cfg_guidance=(B == 2))
    assert y.shape == yo.shape
    assert _rel(y, yo) < 5e-3          # fp16 storage between ops vs the fp32 oracle (GPU bar: same)


def test_refunet_engine_orchestration_matches_oracle(monkeypatch):
    from mikudance_b200 import synth
    from mikudance_b200.engine_ref import RefUNetEngine
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel
    from oracle import refunet_oracle as R
    K.install(monkeypatch)
    cfg = synth.TINY_CONFIG
    sd = synth.synthetic_state_dict(cfg, seed=2, reference_unet=True)
    model = UNet2DConditionModel(block_out_channels=cfg["block_out_channels"],
                                 cross_attention_dim=cfg["cross_attention_dim"])
    model.load_state_dict(sd)
    model = model.half().eval()
    writer = ReferenceAttentionControl(model, mode="write", do_classifier_free_guidance=True,
                                       fusion_blocks="full")
    # 32x32 latents: the deepest level still has 4x4 pixels (InstanceNorm over 1-2 pixels is ill-conditioned
    # for ANY fp16 pipeline: the fp16-emulated reference path itself deviates by 9e-2 there)
    N, h, w = 2, 32, 32
    x, ctx = synth.synthetic_reference_inputs(cfg, N, h, w, lctx=7)
    eng = K.engine_on_cpu(RefUNetEngine, model)
    y, banks = eng.forward_api(x.half(), 0, ctx.half())
    with torch.no_grad():
        yo, bo = R.refunet_forward({k: v.float() for k, v in sd.items()}, cfg, x.half().float(), 0,
                                   ctx.half().float())
    assert set(banks) == set(bo) and len(banks) == 16
    for name in bo:
        assert banks[name].shape == bo[name].shape
        assert _rel(banks[name], bo[name]) < 5e-3, name
    assert _rel(y, yo) < 5e-3
    # the writer contract: forward() appends [N, hw, C] to every hooked block's bank, in pairing order
    model._engine = eng
    monkeypatch.setattr(type(x), "is_cuda", property(lambda self: True), raising=False)
    out = model(x.half(), torch.zeros((), dtype=torch.int64), encoder_hidden_states=ctx.half(),
                return_dict=False)[0]
    assert out.shape == (N, cfg["block_out_channels"][0], h, w)
    for blk, (name, c, ds) in zip(writer._blocks(model), synth.reader_bank_order(cfg)):
        assert len(blk.bank) == 1 and torch.equal(blk.bank[0], banks[name])
        assert blk.bank[0].shape == (N, (h // ds) * (w // ds), c)
    writer.clear()
    assert all(len(b.bank) == 0 for b in writer._blocks(model))


def test_forward_argument_errors(monkeypatch):
    """Error behaviour of the forward call surfaces (raised before any kernel is launched)."""
    from mikudance_b200 import synth
    from mikudance_b200.engine import UNetEngine
    from mikudance_b200.engine_ref import RefUNetEngine
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel
    K.install(monkeypatch)
    cfg = synth.TINY_CONFIG
    m = _unet3d(cfg).half().eval()
    eng = K.engine_on_cpu(UNetEngine, m)
    ctx = torch.zeros(1, 3, cfg["cross_attention_dim"], dtype=torch.float16)
    with pytest.raises(ValueError, match="divisible by 8"):
        eng.forward_api(torch.zeros(1, 4, 2, 12, 8, dtype=torch.float16), 1, ctx)
    with pytest.raises(ValueError, match="temporal_position_encoding_max_len"):
        eng.forward_api(torch.zeros(1, 4, 33, 8, 8, dtype=torch.float16), 1, ctx)        # window > PE table
    with pytest.raises(ValueError, match="batch"):
        eng.forward_api(torch.zeros(3, 4, 2, 8, 8, dtype=torch.float16), 1, ctx.repeat(2, 1, 1))
    with pytest.raises(NotImplementedError, match="per-sample timesteps"):
        eng.forward_api(torch.zeros(2, 4, 2, 8, 8, dtype=torch.float16), torch.tensor([1, 2]), ctx)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 4, 2, 8, 8), 1, ctx, class_labels=torch.zeros(1))
    r = UNet2DConditionModel(block_out_channels=cfg["block_out_channels"],
                             cross_attention_dim=cfg["cross_attention_dim"]).half().eval()
    reng = K.engine_on_cpu(RefUNetEngine, r)
    with pytest.raises(ValueError, match="22|input"):
        reng.forward_api(torch.zeros(1, 4, 16, 16, dtype=torch.float16), 0, ctx)
    with pytest.raises(ValueError, match="more than 1 spatial element"):
        reng.forward_api(torch.zeros(1, 22, 8, 8, dtype=torch.float16), 0, ctx)          # nn.InstanceNorm2d raises too
    with pytest.raises(ValueError, match="divisible by 8"):
        reng.forward_api(torch.zeros(1, 22, 20, 16, dtype=torch.float16), 0, ctx)


def test_midup_fusion_blocks_only_read_mid_and_up_banks(monkeypatch):
    """fusion_blocks="midup" (ReferenceAttentionControl's default): down blocks keep empty banks and fall back to plain
    self-attention (mutual_mix_attention.py:169-172); mid / up blocks add theirs."""
    from mikudance_b200 import synth
    from mikudance_b200.engine import UNetEngine
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from oracle import unet3d_oracle as O
    K.install(monkeypatch)
    cfg = synth.TINY_CONFIG
    sd = synth.synthetic_state_dict(cfg, seed=1)
    model = _unet3d(cfg)
    model.load_state_dict(sd)
    model = model.half().eval()
    B, f, h, w = 2, 2, 8, 8
    x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=5)
    banks = {k: v for k, v in synth.synthetic_banks(cfg, B * f, h, w).items() if not k.startswith("down_blocks")}
    reader = ReferenceAttentionControl(model, mode="read", do_classifier_free_guidance=True, fusion_blocks="midup")
    names = {id(m): n for n, m in model.named_modules()}
    blocks = reader._blocks(model)
    assert len(blocks) == 10 and not any(names[id(b)].startswith("down_blocks") for b in blocks)
    for blk in blocks:
        blk.bank = [banks[names[id(blk)].rsplit(".transformer_blocks", 1)[0]]]
    eng = K.engine_on_cpu(UNetEngine, model)
    y = eng.forward_api(x.half(), 19, ctx.half())
    with torch.no_grad():
        yo = O.unet3d_forward({k: v.float() for k, v in sd.items()}, cfg, x.half().float(), 19, ctx.half().float(),
                              banks=banks, cfg_guidance=True)
    assert _rel(y, yo) < 5e-3
