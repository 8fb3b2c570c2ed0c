"""`src.utils.util` of the reference as imported by scripts/inference_video.py:24 — `get_fps`, `read_frames`,
`save_videos_grid` (reference: src/utils/util.py:51-128).  Host-side media IO, outside the hot path; the reference
uses PyAV, which is not available here, so these go through OpenCV (`cv2`) and PIL with the same signatures and
return types (PIL RGB frames; a frame-rate number; an .mp4 / .gif written to `path`)."""
from __future__ import annotations

import os
from fractions import Fraction
from pathlib import Path
from typing import List

import numpy as np
import torch
from PIL import Image


def seed_everything(seed: int) -> None:
    import random
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed % (2 ** 32))
    random.seed(seed)


def read_frames(video_path: str) -> List[Image.Image]:
    import cv2
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {video_path}")
    frames = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        frames.append(Image.fromarray(np.ascontiguousarray(bgr[:, :, ::-1])))
    cap.release()
    return frames


def get_fps(video_path: str):
    import cv2
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise FileNotFoundError(f"cannot open video {video_path}")
    fps = cap.get(cv2.CAP_PROP_FPS)
    cap.release()
    return Fraction(fps).limit_denominator(1001)          # PyAV's average_rate is a Fraction as well


def save_videos_from_pil(pil_images: List[Image.Image], path: str, fps: int = 8) -> None:
    import cv2
    suffix = Path(path).suffix.lower()
    if os.path.dirname(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
    if suffix == ".mp4":
        w, h = pil_images[0].size
        writer = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
        if not writer.isOpened():
            raise RuntimeError(f"cannot open {path} for writing")
        for im in pil_images:
            writer.write(np.ascontiguousarray(np.asarray(im.convert("RGB"))[:, :, ::-1]))
        writer.release()
    elif suffix == ".gif":
        pil_images[0].save(fp=path, format="GIF", append_images=pil_images[1:], save_all=True,
                           duration=1.0 / fps * 1000, loop=0)
    else:
        raise ValueError("Unsupported file type. Use .mp4 or .gif.")


save_video_from_pil2 = save_videos_from_pil


def save_videos_grid(videos: torch.Tensor, path: str, rescale: bool = False, n_rows: int = 6, fps: int = 8) -> None:
    """videos [b, c, t, h, w] in [0, 1] (or [-1, 1] with rescale): every time step becomes one grid image with up to
    `n_rows` clips per row (2-pixel padding, like torchvision.utils.make_grid's default)."""
    b, c, t, h, w = videos.shape
    pad = 2
    ncol = min(n_rows, b)
    nrow = (b + ncol - 1) // ncol
    frames = []
    v = videos.detach().float().cpu()
    if c == 1:
        v = v.repeat(1, 3, 1, 1, 1)
    for ti in range(t):
        grid = torch.zeros(3, nrow * (h + pad) + pad, ncol * (w + pad) + pad) if b > 1 else None
        if grid is None:
            x = v[0, :, ti]
        else:
            for i in range(b):
                r, col = divmod(i, ncol)
                grid[:, pad + r * (h + pad): pad + r * (h + pad) + h, pad + col * (w + pad): pad + col * (w + pad) + w] = v[i, :, ti]
            x = grid
        if rescale:
            x = (x + 1.0) / 2.0
        arr = (x.clamp(0, 1).permute(1, 2, 0) * 255).numpy().astype(np.uint8)
        frames.append(Image.fromarray(arr))
    save_videos_from_pil(frames, path, fps)
