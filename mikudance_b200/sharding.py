"""Frame sharding of a context window over the GPUs of one box — pure host logic.

Unit of work = one image = (CFG branch, frame) of a window.  Conv, GroupNorm (per frame), spatial /
cross attention and feed-forwards never mix images; only the motion modules' temporal attention mixes
the frames of one CFG branch at one pixel.  Rank r of G owns frames [r*fl, (r+1)*fl) of every window
for both CFG branches.  Inside each motion module the activation is exchanged frame-sharded ->
pixel-sharded (`frames_to_pixels`, one all-to-all), the whole temporal transformer runs on ALL frames
of hw/G pixels with no further communication, and a second all-to-all (`pixels_to_frames`) brings it
back.  The alternative mode all-gathers the temporal K/V rows of all ranks; the temporal-attention
kernel then addresses frame j of batch b in the gathered buffer with `gathered_row()` below (the
formula compiled into csrc/temporal_attn.cu).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


def plan_ranks(rank: int, world: int, do_cfg: bool, cfg_split: bool) -> Dict[str, int]:
    """Which images rank `rank` of `world` evaluates.
    Frame sharding (default): every rank runs both CFG branches of its frames: branch = -1 (both),
    sub_rank = rank, sub_world = world.
    CFG split (`cfg_split`, needs guidance and an even world): the two CFG branches never interact inside
    the UNet (they only meet in the guidance formula, pipeline_mikudance.py:670-674), so ranks
    [0, world/2) take the uncond branch and [world/2, world) the cond branch; frames are sharded over the
    world/2 ranks of a branch and the motion-module exchange stays inside that half-size group (none at
    all at world == 2)."""
    if cfg_split and do_cfg and world >= 2 and world % 2 == 0:
        half = world // 2
        return dict(branch=rank // half, sub_rank=rank % half, sub_world=half)
    return dict(branch=-1, sub_rank=rank, sub_world=world)


def frame_split(length: int, world: int) -> List[int]:
    """Frames of a `length`-frame window per rank: as even as possible, the first length % world ranks carry one
    more (a 30-frame window over 8 GPUs = 6 x 4 + 2 x 3, over 4 GPUs = 2 x 8 + 2 x 7)."""
    if length < world:
        raise ValueError(f"a window of {length} frames cannot be sharded over {world} GPUs (fewer frames than ranks)")
    base, rem = divmod(length, world)
    return [base + 1] * rem + [base] * (world - rem)


def shard_window(window: Sequence[int], rank: int, world: int) -> Tuple[List[int], int]:
    """(frames of `window` owned by `rank`, offset of the first one inside the window)."""
    counts = frame_split(len(window), world)
    lo = sum(counts[:rank])
    return list(window[lo:lo + counts[rank]]), lo


def slice_bank(bank: torch.Tensor, nb: int, window_len: int, rank: int, world: int) -> torch.Tensor:
    """bank [(nb * L), hw, C] for the whole window -> this rank's [(nb * fl), hw, C] (b-major)."""
    counts = frame_split(window_len, world)
    lo, fl = sum(counts[:rank]), counts[rank]
    hw, C = bank.shape[-2], bank.shape[-1]
    return bank.reshape(nb, window_len, hw, C)[:, lo:lo + fl].reshape(nb * fl, hw, C)


def gathered_row(j: int, b: int, px: int, nb: int, fl: int, npix: int) -> int:
    """Row of frame j (window position), batch b, pixel px in the all-gathered [G, nb, fl, npix] K/V
    (even split only: the all-gather mode needs len(window) % world == 0)."""
    g, l = divmod(j, fl)
    return ((g * nb + b) * fl + l) * npix + px


def pixels_per_rank(hw: int, world: int) -> int:
    """Pixels each rank owns in the pixel-sharded layout (the last rank's tail is padding)."""
    return (hw + world - 1) // world


def frames_to_pixels(send: torch.Tensor, out: torch.Tensor, counts: Sequence[int], pp: int, rank: int,
                     group) -> torch.Tensor:
    """One CFG branch, frame-sharded -> pixel-sharded, ONE all-to-all and no layout copy on either side.
    send [G * fl * pp, C]: this rank's fl = counts[rank] frames in the exchange layout [G(dst), fl, pp, C] that
         mdk_groupnorm_f16 (`chunks=(G, pp)`) writes directly: block d holds pixels [d*pp, (d+1)*pp) of every frame;
    out  [F * pp, C], F = sum(counts): ALL frames (window order) of this rank's pp pixels — rank s's block
         [counts[s], pp, C] lands at frame offset sum(counts[:s]), so the result is frame-major as it is."""
    import torch.distributed as dist
    world = len(counts)
    fl = counts[rank]
    dist.all_to_all_single(out, send, output_split_sizes=[c * pp for c in counts],
                           input_split_sizes=[fl * pp] * world, group=group)
    return out


def pixels_to_frames(h: torch.Tensor, back: torch.Tensor, counts: Sequence[int], pp: int, rank: int,
                     group) -> torch.Tensor:
    """Inverse exchange of one CFG branch: h [F * pp, C] (all frames of this rank's pixels, frame-major: the rows
    of rank d's frames are contiguous) -> back [G(chunk), fl, pp, C] (this rank's frames, one block per pixel
    chunk); `ops.unshard` puts the chunks back side by side."""
    import torch.distributed as dist
    world = len(counts)
    fl = counts[rank]
    dist.all_to_all_single(back, h, output_split_sizes=[fl * pp] * world,
                           input_split_sizes=[c * pp for c in counts], group=group)
    return back
