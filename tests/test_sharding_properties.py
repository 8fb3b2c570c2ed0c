"""CPU, property-based (hypothesis): the pure partitioning functions of mikudance_b200.sharding and the window
scheduler — every frame owned exactly once, bank slices line up with frame shards, the gathered-row formula is a
bijection, CFG-split plans cover both branches, window lists cover every frame."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from mikudance_b200.context import uniform
from mikudance_b200.sharding import (frame_split, gathered_row, pixels_per_rank, plan_ranks, shard_window,
                                     slice_bank)


@settings(max_examples=80, deadline=None)
@given(world=st.sampled_from([1, 2, 3, 4, 8]), extra=st.integers(0, 20), start=st.integers(0, 40))
def test_shards_partition_the_window(world, extra, start):
    L = world + extra                                  # any length >= world, divisible or not
    window = [(start + i) % 64 for i in range(L)]
    counts = frame_split(L, world)
    assert sum(counts) == L and max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)
    seen = []
    for r in range(world):
        mine, lo = shard_window(window, r, world)
        assert lo == sum(counts[:r]) and mine == window[lo:lo + counts[r]]
        seen += mine
    assert seen == window


@settings(max_examples=40, deadline=None)
@given(world=st.sampled_from([1, 2, 4]), extra=st.integers(0, 7), nb=st.sampled_from([1, 2]), hw=st.integers(1, 5))
def test_bank_slices_follow_the_frame_shards(world, extra, nb, hw):
    L = world + extra
    counts = frame_split(L, world)
    bank = torch.arange(nb * L * hw * 2, dtype=torch.float32).reshape(nb * L, hw, 2)
    parts = [slice_bank(bank, nb, L, r, world).reshape(nb, counts[r], hw, 2) for r in range(world)]
    assert torch.equal(torch.cat(parts, dim=1).reshape(nb * L, hw, 2), bank)


@settings(max_examples=40, deadline=None)
@given(world=st.sampled_from([1, 2, 4, 8]), fl=st.integers(1, 4), nb=st.sampled_from([1, 2]), npix=st.integers(1, 6))
def test_gathered_row_is_a_bijection(world, fl, nb, npix):
    rows = {gathered_row(j, b, px, nb, fl, npix) for j in range(world * fl) for b in range(nb) for px in range(npix)}
    assert rows == set(range(world * nb * fl * npix))
    # rank g's block holds exactly its own frames, batch-major, as all_gather_into_tensor lays them out
    for g in range(world):
        lo = g * nb * fl * npix
        assert gathered_row(g * fl, 0, 0, nb, fl, npix) == lo
        assert gathered_row(g * fl + fl - 1, nb - 1, npix - 1, nb, fl, npix) == lo + nb * fl * npix - 1


@settings(max_examples=40, deadline=None)
@given(world=st.integers(1, 16), hw=st.integers(1, 300))
def test_pixel_shards_cover_every_pixel(world, hw):
    pp = pixels_per_rank(hw, world)
    assert pp * world >= hw and (pp - 1) * world < hw


@settings(max_examples=60, deadline=None)
@given(world=st.integers(1, 16), do_cfg=st.booleans(), split=st.booleans())
def test_plan_ranks_covers_every_image_once(world, do_cfg, split):
    plans = [plan_ranks(r, world, do_cfg, split) for r in range(world)]
    if split and do_cfg and world >= 2 and world % 2 == 0:
        half = world // 2
        for b in (0, 1):
            assert sorted(p["sub_rank"] for p in plans if p["branch"] == b) == list(range(half))
        assert all(p["sub_world"] == half for p in plans)
    else:
        assert all(p == dict(branch=-1, sub_rank=r, sub_world=world) for r, p in enumerate(plans))


@settings(max_examples=60, deadline=None)
@given(frames=st.integers(1, 120), ctx=st.integers(2, 40), stride=st.integers(1, 3), data=st.data())
def test_windows_cover_every_frame(frames, ctx, stride, data):
    overlap = data.draw(st.integers(0, max(0, ctx - 1)))
    wins = list(uniform(0, 20, frames, ctx, stride, overlap))
    assert wins, "at least one window"
    covered = set()
    for w in wins:
        assert len(w) == min(ctx, frames) and all(0 <= f < frames for f in w)
        covered.update(w)
    assert covered == set(range(frames))
