"""CPU, world_size 2 and 4 over gloo: the PRODUCT's frame-sharded UNet forward (`UNetEngine.run` with a
process group: all-to-all frame<->pixel exchange per motion module, or the all-gather of the temporal
K/V rows) against the same engine on one process, with every kernel call routed to the CPU statement of
its C-ABI contract (tests/ops_contract_cpu.py).  This covers the N>1 host path end to end — PE offsets,
bank slicing, uncond/cond split per shard, padded pixel shards — without a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q, mode, h, w, F_):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ops_contract_cpu as K
        from mikudance_b200 import ops, synth
        from mikudance_b200.engine import UNetEngine
        from mikudance_b200.sharding import shard_window, slice_bank
        from mikudance_b200.unet_3d import UNet3DConditionModel
        for n in K._NAMES:
            setattr(ops, n, getattr(K, n))
        torch.set_num_threads(2)
        cfg = synth.TINY_CONFIG
        m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"],
                                 cross_attention_dim=cfg["cross_attention_dim"], use_inflated_groupnorm=True,
                                 use_motion_module=True, motion_module_mid_block=True, motion_module_type="Vanilla",
                                 unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
        m.load_state_dict(synth.synthetic_state_dict(cfg, seed=0))
        m = m.half().eval()
        nb, lctx = 2, 5
        x, ctx = synth.synthetic_inputs(cfg, nb, F_, h, w, lctx=lctx)
        banks = synth.synthetic_banks(cfg, nb * F_, h, w)
        window = list(range(F_))
        x16, c16 = x.half(), ctx.half()
        # one process, whole window
        eng1 = K.engine_on_cpu(UNetEngine, m)
        eng1.t_dev.fill_(499)
        full_in = ops.latents_to_nhwc(x16[:1].contiguous(), b=nb, frame_idx=None, fl=F_, cpad=eng1.cin_pad)
        full = eng1.run(full_in, nb, F_, h, w, c16, {k: v.reshape(-1, v.shape[-1]) for k, v in banks.items()},
                        n_uncond=F_).view(nb, F_, h * w, -1)
        # this rank's shard
        mine, lo = shard_window(window, rank, world)
        fl = len(mine)
        eng = K.engine_on_cpu(UNetEngine, m)
        eng.t_dev.fill_(499)
        eng.shard_mode = mode
        eng.set_process_group(dist.group.WORLD, rank, world)
        idx = torch.tensor(mine, dtype=torch.int32)
        loc_in = ops.latents_to_nhwc(x16[:1].contiguous(), b=nb, frame_idx=idx, fl=fl, cpad=eng.cin_pad)
        loc_banks = {k: slice_bank(v, nb, F_, rank, world).reshape(-1, v.shape[-1]).contiguous()
                     for k, v in banks.items()}
        got = eng.run(loc_in, nb, fl, h, w, c16, loc_banks, n_uncond=fl, f_off=lo, f_total=F_)
        want = full[:, lo:lo + fl].reshape(nb * fl * h * w, -1)
        rel = ((got.float() - want.float()).norm() / want.float().norm()).item()
        q.put((rank, rel, bool(torch.isfinite(got.float()).all())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode,h,w,F_", [(2, "a2a", 8, 8, 4), (2, "allgather", 8, 8, 4),
                                                (4, "a2a", 16, 8, 4), (2, "a2a", 8, 8, 5), (4, "a2a", 8, 8, 6)])
def test_sharded_engine_equals_single_process(world, mode, h, w, F_):
    port = 33500 + (os.getpid() % 2000) + world
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, h, w, F_)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in res) == list(range(world))
    for rank, rel, finite in res:
        # identical arithmetic per image (float64 contract math): the shard equals the single-process rows
        assert finite and rel < 1e-6, (rank, rel)
