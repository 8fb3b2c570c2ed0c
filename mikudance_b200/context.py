"""Sliding-window frame scheduler of the denoising loop (the reference's
src/pipelines/context.py:7-76): integer index lists, host side, bit-exact with the reference."""
from __future__ import annotations

import math
from typing import Callable, Iterator, List, Optional


def ordered_halving(val: int) -> float:
    """Van-der-Corput style fraction: reverse the 64 bits of `val`, read as a fraction of 2^64."""
    rev = 0
    v = int(val)
    for _ in range(64):
        rev = (rev << 1) | (v & 1)
        v >>= 1
    return rev / (1 << 64)


def uniform(step: int = 0, num_steps: Optional[int] = None, num_frames: int = 0,
            context_size: Optional[int] = None, context_stride: int = 3, context_overlap: int = 4,
            closed_loop: bool = True) -> Iterator[List[int]]:
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    levels = min(context_stride, int(math.ceil(math.log2(num_frames / context_size))) + 1)
    frac = ordered_halving(step)
    for level in range(levels):
        cstep = 1 << level
        pad = int(round(num_frames * frac))
        first = int(frac * cstep) + pad
        last = num_frames + pad + (0 if closed_loop else -context_overlap)
        hop = context_size * cstep - context_overlap
        for j in range(first, last, hop):
            yield [e % num_frames for e in range(j, j + context_size * cstep, cstep)]


def get_context_scheduler(name: str) -> Callable:
    if name == "uniform":
        return uniform
    raise ValueError(f"Unknown context_overlap policy {name}")


def get_total_steps(scheduler, timesteps, num_steps=None, num_frames=0, context_size=None,
                    context_stride=3, context_overlap=4, closed_loop=True) -> int:
    return sum(len(list(scheduler(i, num_steps, num_frames, context_size, context_stride,
                                  context_overlap))) for i in range(len(timesteps)))
