"""ORACLE (test infrastructure).  Sliding-window frame scheduler restated from
src/pipelines/context.py:7-42 of the reference (`uniform`, `ordered_halving`).  The reference file
imports untouched, so tests/test_oracle_vs_reference.py compares the two directly and
tests/golden/context_windows.json pins outputs generated from the reference itself."""
import math
from typing import List


def ordered_halving(val: int) -> float:
    """bit-reverse a 64-bit integer and read it as a fraction in [0, 1) (context.py:7-12)."""
    return int(f"{val:064b}"[::-1], 2) / (1 << 64)


def uniform(step: int, num_steps, num_frames: int, context_size: int, context_stride: int = 3,
            context_overlap: int = 4, closed_loop: bool = True) -> List[List[int]]:
    """context.py:15-42."""
    if num_frames <= context_size:
        return [list(range(num_frames))]
    context_stride = min(context_stride, int(math.ceil(math.log2(num_frames / context_size))) + 1)
    out = []
    for k in range(context_stride):
        context_step = 1 << k
        pad = int(round(num_frames * ordered_halving(step)))
        start = int(ordered_halving(step) * context_step) + pad
        stop = num_frames + pad + (0 if closed_loop else -context_overlap)
        stride = context_size * context_step - context_overlap
        for j in range(start, stop, stride):
            out.append([e % num_frames for e in range(j, j + context_size * context_step, context_step)])
    return out
