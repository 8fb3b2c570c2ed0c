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


def shard_window(window: Sequence[int], rank: int, world: int) -> Tuple[List[int], int]:
    """(frames of `window` owned by `rank`, offset of the first one inside the window)."""
    L = len(window)
    if L % world != 0:
        raise ValueError(f"a window of {L} frames cannot be split evenly over {world} GPUs; "
                         "choose context_frames divisible by the GPU count")
    fl = L // world
    lo = rank * fl
    return list(window[lo:lo + fl]), lo


def slice_bank(bank: torch.Tensor, nb: int, window_len: int, rank: int, world: int) -> torch.Tensor:
    """bank [(nb * L), hw, C] for the whole window -> this rank's [(nb * fl), hw, C] (b-major)."""
    fl = window_len // world
    hw, C = bank.shape[-2], bank.shape[-1]
    return bank.reshape(nb, window_len, hw, C)[:, rank * fl:(rank + 1) * fl].reshape(nb * fl, hw, C)


def gathered_row(j: int, b: int, px: int, nb: int, fl: int, npix: int) -> int:
    """Row of frame j (window position), batch b, pixel px in the all-gathered [G, nb, fl, npix] K/V."""
    g, l = divmod(j, fl)
    return ((g * nb + b) * fl + l) * npix + px


def pixels_per_rank(hw: int, world: int) -> int:
    """Pixels each rank owns in the pixel-sharded layout (the last rank's tail is zero padding)."""
    return (hw + world - 1) // world


def frames_to_pixels(x: torch.Tensor, nb: int, fl: int, hw: int, world: int, group) -> torch.Tensor:
    """x [(nb fl) hw, C], this rank's fl frames of every pixel  ->  [(nb F) pp, C] with F = world*fl:
    ALL frames (window order: source rank major) of this rank's pp = ceil(hw / world) pixels.
    Pixels >= hw (only when hw % world != 0) are zero rows.  ONE all-to-all for both CFG branches."""
    import torch.distributed as dist
    C = x.shape[-1]
    pp = pixels_per_rank(hw, world)
    xv = x.view(nb, fl, hw, C)
    if pp * world != hw:
        pad = torch.zeros((nb, fl, pp * world - hw, C), dtype=x.dtype, device=x.device)
        xv = torch.cat([xv, pad], dim=2)
    send = xv.view(nb, fl, world, pp, C).permute(2, 0, 1, 3, 4).contiguous()   # [G(dst), nb, fl, pp, C]
    recv = torch.empty_like(send)                                              # [G(src), nb, fl, pp, C]
    dist.all_to_all_single(recv, send, group=group)
    return recv.permute(1, 0, 2, 3, 4).reshape(nb * world * fl * pp, C)        # [nb, G(src), fl, pp, C]


def pixels_to_frames(h: torch.Tensor, nb: int, fl: int, hw: int, world: int, group) -> torch.Tensor:
    """Inverse of frames_to_pixels: h [(nb F) pp, C] -> [(nb fl) hw, C]."""
    import torch.distributed as dist
    C = h.shape[-1]
    pp = pixels_per_rank(hw, world)
    send = h.view(nb, world, fl, pp, C).permute(1, 0, 2, 3, 4).contiguous()    # [G(dst frames), nb, fl, pp, C]
    back = torch.empty_like(send)                                              # [G(pixel chunk), nb, fl, pp, C]
    dist.all_to_all_single(back, send, group=group)
    out = back.permute(1, 2, 0, 3, 4).reshape(nb, fl, world * pp, C)
    if pp * world != hw:
        out = out[:, :, :hw]
    return out.reshape(nb * fl * hw, C)
