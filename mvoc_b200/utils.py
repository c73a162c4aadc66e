"""Host-side helpers mirroring i2vgen-xl/utils.py for the pieces that fix data layout on the hot path:
``mask_preprocess`` (utils.py:92-154), ``load_ddim_latents_at_t`` / ``load_ddim_latents_at_T``
(utils.py:31-45), ``seed_everything`` (utils.py:23-28).  PIL / OpenCV run on the host once per run (setup,
not the hot path); the results are the ``(float [1,4,T,h,w], bool [1,4,T,h,w])`` pairs that
``register_time_all`` pushes to the hooks (pipelines/pipeline_i2vgen_xl.py:1588-1599, :1684-1685).
"""
from __future__ import annotations

import glob
import os
import random
from typing import Tuple

import numpy as np
import torch

from .pipeline import load_ddim_latents_at_t  # noqa: F401  (re-export, utils.py:31-36)


def seed_everything(seed: int) -> None:
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    random.seed(seed)
    np.random.seed(seed)


def load_ddim_latents_at_T(ddim_latents_path: str) -> torch.Tensor:
    """utils.py:39-45 — the noisiest stored level."""
    ts = [int(os.path.basename(p).split("_")[-1].split(".")[0])
          for p in glob.glob(os.path.join(ddim_latents_path, "ddim_latents_*.pt"))]
    if not ts:
        raise FileNotFoundError(f"no ddim_latents_*.pt under {ddim_latents_path}")
    return torch.load(os.path.join(ddim_latents_path, f"ddim_latents_{max(ts)}.pt"), map_location="cpu")


def _one_mask(path: str, downscale: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """One PNG -> (float [h,w] in [0,1] with 255 levels, bool [h,w]) — utils.py:93-109 / :123-138."""
    import cv2 as cv
    from PIL import Image

    mask = Image.open(path).convert("L")
    w, h = mask.size
    mask = mask.resize((w // downscale, h // downscale))           # PIL default resample (bicubic for "L")
    arr = np.asarray(mask)
    _, binary = cv.threshold(arr, 10, 255, cv.THRESH_BINARY)       # > 10 -> 255
    mf = torch.from_numpy(arr.copy()).to(torch.float32).div_(255.0)
    mb = torch.from_numpy(binary.copy()).to(torch.float32).div_(255.0).to(torch.bool)
    return mf, mb


def mask_preprocess(mask: str, device, dtype, batch_size: int, channel: int, frames: int, downscale: int = 8):
    """utils.py:148-154.  `mask` is a PNG (static object, repeated over the frames, :92-109) or a folder of
    numbered PNGs (one per frame, sorted numerically, truncated to `frames`, :112-145).
    Returns (float mask in `dtype`, bool mask), both [batch_size, channel, T, H/downscale, W/downscale]."""
    if os.path.isdir(mask):
        paths = glob.glob(os.path.join(mask, "*.png"))
        paths.sort(key=lambda p: int(os.path.basename(p).split(".")[0]))
        if len(paths) != frames:
            paths = paths[:frames]
        pairs = [_one_mask(p, downscale) for p in paths]
        mf = torch.stack([a for a, _ in pairs], dim=0)
        mb = torch.stack([b for _, b in pairs], dim=0)
    else:
        a, b = _one_mask(mask, downscale)
        mf = a[None].expand(frames, *a.shape)
        mb = b[None].expand(frames, *b.shape)
    mf = mf.to(device=device, dtype=dtype)[None, None].expand(batch_size, channel, *mf.shape).contiguous()
    mb = mb.to(device=device)[None, None].expand(batch_size, channel, *mb.shape).contiguous()
    return mf, mb
