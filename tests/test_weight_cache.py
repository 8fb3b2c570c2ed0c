"""CPU: the packed fp16 weight cache (mikudance_b200/weight_cache.py).
A cache hit / a packed checkpoint must reproduce the freshly packed engine tensor for tensor and the forward bit for
bit (kernel calls routed to the CPU statement of the C-ABI contract, tests/ops_contract_cpu.py); a changed weight,
a changed layout version or a different GEGLU panel width must miss."""
import os

import pytest
import torch

import ops_contract_cpu as K


def _unet3d(cfg, seed=1):
    from mikudance_b200 import synth
    from mikudance_b200.unet_3d import UNet3DConditionModel
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"], cross_attention_dim=cfg["cross_attention_dim"],
                             use_inflated_groupnorm=True, use_motion_module=True, motion_module_mid_block=True,
                             motion_module_type="Vanilla", unet_use_cross_frame_attention=False,
                             unet_use_temporal_attention=False)
    m.load_state_dict(synth.synthetic_state_dict(cfg, seed=seed))
    return m.half().eval()


def _packed_tensors(eng):
    from mikudance_b200 import weight_cache
    return weight_cache.export_packed(eng)


def _same_packed(a, b):
    ta, sa = _packed_tensors(a)
    tb, sb = _packed_tensors(b)
    assert sa == sb and set(ta) == set(tb) and len(ta) > 100
    scratch = ("te_scratch", "temb_vec")                     # torch.empty buffers: contents are not weights
    for k in ta:
        assert ta[k].dtype == tb[k].dtype and ta[k].shape == tb[k].shape, k
        if k not in scratch:
            assert torch.equal(ta[k], tb[k]), k


def test_weights_key_is_content_and_position_sensitive():
    from mikudance_b200 import synth, weight_cache
    cfg = synth.TINY_CONFIG
    m = _unet3d(cfg)
    k0 = weight_cache.weights_key(m, "t")
    assert k0 == weight_cache.weights_key(_unet3d(cfg), "t") and len(k0) == 32
    assert weight_cache.weights_key(m, "u") != k0                     # layout tag is part of the key
    w = m.down_blocks[1].resnets[0].conv1.weight
    with torch.no_grad():
        a, b = w.view(-1)[3].clone(), w.view(-1)[5000].clone()
        assert a != b
        w.view(-1)[3], w.view(-1)[5000] = b, a                        # a swap: same multiset of values
    k1 = weight_cache.weights_key(m, "t")
    assert k1 != k0
    with torch.no_grad():
        w.view(-1)[3], w.view(-1)[5000] = a, b
    assert weight_cache.weights_key(m, "t") == k0
    assert weight_cache.weights_key(_unet3d(cfg, seed=2), "t") != k0


@pytest.mark.parametrize("which", ["unet3d", "refunet"])
def test_transparent_cache_hit_reproduces_the_packed_engine(monkeypatch, tmp_path, which):
    from mikudance_b200 import synth, weight_cache
    from mikudance_b200.engine import UNetEngine
    from mikudance_b200.engine_ref import RefUNetEngine
    from mikudance_b200.unet_2d_ref import UNet2DConditionModel
    K.install(monkeypatch)
    cfg = synth.TINY_CONFIG
    if which == "unet3d":
        cls, model = UNetEngine, _unet3d(cfg)
    else:
        model = UNet2DConditionModel(block_out_channels=cfg["block_out_channels"], cross_attention_dim=cfg["cross_attention_dim"])
        model.load_state_dict(synth.synthetic_state_dict(cfg, seed=2, reference_unet=True))
        cls, model = RefUNetEngine, model.half().eval()
    plain = K.engine_on_cpu(cls, model)
    assert plain.weight_cache is None                                   # MDK_WEIGHT_CACHE unset: no file I/O
    monkeypatch.setenv("MDK_WEIGHT_CACHE", str(tmp_path))
    first = K.engine_on_cpu(cls, model)
    files = sorted(os.listdir(tmp_path))
    assert first.weight_cache == "miss" and len(files) == 1 and files[0].startswith("mdk-packed-")
    calls = []
    monkeypatch.setattr(cls, "_pack", lambda self: calls.append(1))     # a hit must not pack
    second = K.engine_on_cpu(cls, model)
    assert second.weight_cache == "hit" and not calls
    _same_packed(plain, second)
    monkeypatch.undo()
    K.install(monkeypatch)
    monkeypatch.setenv("MDK_WEIGHT_CACHE", str(tmp_path))
    # a changed weight -> another key -> miss, second file
    with torch.no_grad():
        model.conv_in.weight.view(-1)[0] += 1.0
    third = K.engine_on_cpu(cls, model)
    assert third.weight_cache == "miss" and len(os.listdir(tmp_path)) == 2
    # same weights, another kernel layout version -> another key; a file whose header disagrees is "stale"
    monkeypatch.setattr(weight_cache, "LAYOUT_VERSION", weight_cache.LAYOUT_VERSION + 1)
    assert K.engine_on_cpu(cls, model).weight_cache == "miss" and len(os.listdir(tmp_path)) == 3


def test_stale_header_is_refused(monkeypatch, tmp_path):
    from mikudance_b200 import synth, weight_cache
    from mikudance_b200.engine import UNetEngine
    K.install(monkeypatch)
    model = _unet3d(synth.TINY_CONFIG)
    eng = K.engine_on_cpu(UNetEngine, model)
    path = str(tmp_path / "x.safetensors")
    weight_cache.write_file(eng, path, "k")
    fresh = UNetEngine.__new__(UNetEngine)
    fresh.__dict__.update(model=model, dev=torch.device("cpu"), plan=eng.plan, geglu_block=eng.geglu_block)
    assert weight_cache.read_file(fresh, path, "other-key") is False and not hasattr(fresh, "down")
    fresh.geglu_block = eng.geglu_block * 2                              # another build's GEGLU panel width
    assert weight_cache.read_file(fresh, path, "k") is False
    fresh.geglu_block = eng.geglu_block
    assert weight_cache.read_file(fresh, path, "k") is True and len(fresh.down) == 4
    with pytest.raises(RuntimeError, match="not a packed checkpoint"):
        weight_cache.read_ctor(path)                                     # cache entries carry no ctor kwargs


def test_packed_checkpoint_roundtrip_forward_bit_exact(monkeypatch, tmp_path):
    from mikudance_b200 import synth
    from mikudance_b200.engine import UNetEngine
    from mikudance_b200.reference_control import ReferenceAttentionControl
    from mikudance_b200.unet_3d import UNet3DConditionModel
    K.install(monkeypatch)
    cfg = synth.TINY_CONFIG
    model = _unet3d(cfg)
    model._engine = K.engine_on_cpu(UNetEngine, model)
    path = tmp_path / "denoising_unet.packed.safetensors"
    model.save_packed(path)
    # product entry point: CUDA only
    with pytest.raises(RuntimeError, match="no CPU path"):
        UNet3DConditionModel.from_packed(path, device="cpu")
    packed = UNet3DConditionModel._from_packed(path, torch.device("cpu"))
    assert all(p.device.type == "meta" for p in packed.parameters())     # no second copy of the weights
    assert packed.dtype == torch.float16 and packed.device == torch.device("cpu")
    import json
    assert json.dumps(vars(packed.config), sort_keys=True) == json.dumps(vars(model.config), sort_keys=True)   # tuples == lists
    _same_packed(model._engine, packed._engine)
    for call in (lambda: packed.half(), lambda: packed.state_dict(), lambda: packed.load_state_dict({})):
        with pytest.raises(RuntimeError, match="from_packed"):
            call()
    B, f, h, w = 2, 3, 8, 8
    x, ctx = synth.synthetic_inputs(cfg, B, f, h, w, lctx=5)
    banks = synth.synthetic_banks(cfg, B * f, h, w)
    outs = []
    for m in (model, packed):
        ReferenceAttentionControl(m, mode="read", do_classifier_free_guidance=True, fusion_blocks="full")
        for blk, (name, c, ds) in zip(m.spatial_blocks(), synth.reader_bank_order(cfg)):
            blk.bank = [banks[name]]
        outs.append(m._engine.forward_api(x.half(), 499, ctx.half()))
    assert torch.equal(outs[0], outs[1])


def _rank_worker(rank, world, port, cache_dir, q):
    import sys
    import torch.distributed as dist
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), MDK_WEIGHT_CACHE=cache_dir)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mikudance_b200 import ops, synth
        from mikudance_b200.engine import UNetEngine
        model = _unet3d(synth.TINY_CONFIG)
        dist.barrier()                                   # both ranks reach the cache at the same time
        eng = K.engine_on_cpu(UNetEngine, model)
        dist.barrier()
        again = K.engine_on_cpu(UNetEngine, model)
        q.put((rank, eng.weight_cache, again.weight_cache, float(again.temb_w.double().sum())))
    finally:
        dist.destroy_process_group()


def test_ranks_sharing_one_cache_directory_gloo(tmp_path):
    """One process per GPU: every rank may miss and write the same entry concurrently (temp file + atomic rename);
    afterwards there is exactly one file and every rank hits it."""
    import torch.multiprocessing as mp
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[2] for r in res] == ["hit", "hit"] and all(r[1] in ("miss", "hit") for r in res)
    assert res[0][3] == res[1][3]
    files = os.listdir(tmp_path)
    assert len(files) == 1 and files[0].startswith("mdk-packed-") and ".tmp." not in files[0]
