"""CPU, world_size 2 over gloo: the frame-sharding host logic and the all-gather layout the
temporal-attention kernel assumes.  The arithmetic here is the oracle's (this is a
test of partitioning semantics, not of the CUDA kernels)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mikudance_b200 import synth
        from mikudance_b200.sharding import gathered_row, shard_window, slice_bank
        from oracle import unet3d_oracle as O
        torch.manual_seed(0)
        cfg = synth.TINY_CONFIG
        sd = {k: v.float() for k, v in synth.synthetic_state_dict(cfg, seed=0).items()}
        nb, L, H, W, C = 2, 6, 4, 4, 64
        window = [4, 5, 0, 1, 2, 3]
        x = torch.randn(nb * L, C, H, W)                      # (b f) c h w, whole window
        name = "down_blocks.0.motion_modules.0"
        full = O.motion_module(sd, name, x, L, 8, 32)
        mine, lo = shard_window(window, rank, world)
        fl = len(mine)
        assert mine == window[rank * fl:(rank + 1) * fl] and lo == rank * fl
        # --- sharded evaluation: local rows + all-gathered K/V source, PE offset = position in window
        tt = name + ".temporal_transformer"
        xl = x.reshape(nb, L, C, H, W)[:, lo:lo + fl].reshape(nb * fl, C, H, W)
        h = O._gn(sd, tt + ".norm", xl, 32, 1e-6).permute(0, 2, 3, 1).reshape(nb * fl, H * W, C)
        h = O._lin(sd, tt + ".proj_in", h)
        blk = tt + ".transformer_blocks.0"
        for a in range(2):
            n = O._ln(sd, blk + f".norms.{a}", h)                       # [(nb fl), hw, C] local
            gathered = torch.empty(world * nb * fl * H * W, C)
            dist.all_gather_into_tensor(gathered, n.reshape(-1, C).contiguous())
            # rebuild the full-window sequence from the gathered buffer with the kernel's formula
            seq = torch.empty(nb, H * W, L, C)
            for b in range(nb):
                for j in range(L):
                    for px in range(H * W):
                        seq[b, px, j] = gathered[gathered_row(j, b, px, nb, fl, H * W)]
            seq = seq.reshape(nb * H * W, L, C)
            pe = sd[blk + f".attention_blocks.{a}.pos_encoder.pe"].float()
            q_in = (seq + pe[:, :L])[:, lo:lo + fl]                      # local queries, global positions
            o = O._attention(sd, blk + f".attention_blocks.{a}", q_in, seq, 8)
            o = o.reshape(nb, H * W, fl, C).permute(0, 2, 1, 3).reshape(nb * fl, H * W, C)
            h = o + h
        h = O._feed_forward(sd, blk + ".ff", O._ln(sd, blk + ".ff_norm", h)) + h
        h = O._lin(sd, tt + ".proj_out", h).reshape(nb * fl, H, W, C).permute(0, 3, 1, 2) + xl
        want = full.reshape(nb, L, C, H, W)[:, lo:lo + fl].reshape(nb * fl, C, H, W)
        err = (h - want).abs().max().item()
        # banks are sliced the same way
        bank = torch.arange(nb * L * 3 * 2, dtype=torch.float32).reshape(nb * L, 3, 2)
        sl = slice_bank(bank, nb, L, rank, world)
        ok_bank = torch.equal(sl, bank.reshape(nb, L, 3, 2)[:, lo:lo + fl].reshape(nb * fl, 3, 2))
        # window accumulators summed over ranks == unsharded accumulate
        acc = torch.zeros(L)
        acc[lo:lo + fl] = 1.0
        dist.all_reduce(acc)
        q.put((rank, err, ok_bank, bool((acc == 1).all())))
    finally:
        dist.destroy_process_group()


def test_frame_sharding_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ok_bank, ok_acc in res:
        assert err < 1e-4, (rank, err)
        assert ok_bank and ok_acc


def test_uneven_window_split():
    """the pipeline's default 30-frame windows over 4 / 8 GPUs (8 + 8 + 7 + 7, 6 x 4 + 2 x 3)."""
    from mikudance_b200.sharding import frame_split, shard_window
    assert frame_split(30, 4) == [8, 8, 7, 7] and frame_split(30, 8) == [4, 4, 4, 4, 4, 4, 3, 3]
    assert shard_window(list(range(30)), 2, 4) == (list(range(16, 23)), 16)
    assert shard_window(list(range(32)), 3, 4) == (list(range(24, 32)), 24)
    with pytest.raises(ValueError):
        shard_window(list(range(3)), 0, 4)          # fewer frames than ranks


def _worker_a2a(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ops_contract_cpu as K
        from mikudance_b200.sharding import frame_split, frames_to_pixels, pixels_per_rank, pixels_to_frames
        res = []
        for hw, F_ in ((6, 3 * world), (5, 3 * world + 1), (1, 2 * world + world - 1)):   # even / uneven frame splits
            C = 8
            counts = frame_split(F_, world)
            lo, fl = sum(counts[:rank]), counts[rank]
            # value encodes (frame, pixel, channel) so any misplaced row is visible
            full = (torch.arange(F_ * hw * C, dtype=torch.float32).reshape(F_, hw, C) % 2000.0 + 1.0).half()
            mine = full[lo:lo + fl].reshape(fl * hw, C).contiguous()
            pp = pixels_per_rank(hw, world)
            # the send buffer exactly as the GroupNorm kernel lays it out (identity affine, checked separately):
            send = torch.zeros(fl, world * pp, C, dtype=torch.float16)
            send[:, :hw] = mine.view(fl, hw, C)
            send = send.view(fl, world, pp, C).permute(1, 0, 2, 3).reshape(world * fl * pp, C).contiguous()
            got = frames_to_pixels(send, torch.empty(F_ * pp, C, dtype=torch.float16), counts, pp, rank, None)
            want = torch.zeros(F_, pp, C, dtype=torch.float16)
            plo, phi = rank * pp, min((rank + 1) * pp, hw)
            if phi > plo:
                want[:, :phi - plo] = full[:, plo:phi]
            ok_fwd = torch.equal(got.view(F_, pp, C), want)
            back = pixels_to_frames(got, torch.empty(world * fl * pp, C, dtype=torch.float16), counts, pp, rank, None)
            ok_rt = torch.equal(K.unshard(back, nimg=fl, hw=hw, chunk_pix=pp), mine)
            res.append((hw, ok_fwd, ok_rt))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_frame_pixel_exchange_gloo(world):
    """frames_to_pixels / pixels_to_frames (the two all-to-alls of a frame-sharded motion module):
    every rank ends up with ALL frames of its pixel slice in window order, and the round trip is the
    identity — including pixel counts that do not divide by the GPU count."""
    port = 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_a2a, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in out) == list(range(world))
    for _, res in out:
        for hw, ok_fwd, ok_rt in res:
            assert ok_fwd, f"frames_to_pixels misplaced rows (hw={hw})"
            assert ok_rt, f"round trip is not the identity (hw={hw})"
