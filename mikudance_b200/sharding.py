"""Frame sharding of a context window over the GPUs of one box — pure host logic.

Unit of work = one image = (CFG branch, frame) of a window.  Conv, GroupNorm (per frame), spatial /
cross attention and feed-forwards never mix images; only the motion modules' temporal attention mixes
the frames of one CFG branch at one pixel.  Rank r of G owns frames [r*fl, (r+1)*fl) of every window
for both CFG branches; inside each motion module the K/V rows of all ranks are all-gathered (NCCL) and
the temporal-attention kernel addresses frame j of batch b in the gathered buffer with
`gathered_row()` below (the formula compiled into csrc/temporal_attn.cu).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


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
