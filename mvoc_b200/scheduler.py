"""Host-side DDIM schedule for the step loops (scalars only; the update itself is the
``mvoc_cfg_ddim_step`` / ``mvoc_ddim_inverse_step`` kernel).

Mirrors the i2vgen-xl scheduler configuration the reference loads with
``DDIMScheduler.from_pretrained`` (composite.py:82-85, inverse.py:122-131): cosine
(``squaredcos_cap_v2``) betas with zero-terminal-SNR rescale, ``leading`` spacing, ``steps_offset=1``,
v-prediction, eta = 0.  Table arithmetic is float32 like diffusers'.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

NUM_TRAIN_TIMESTEPS = 1000


def _alphas_cumprod() -> np.ndarray:
    """float32 torch arithmetic in the same order as diffusers (betas_for_alpha_bar ->
    rescale_zero_terminal_snr -> cumprod), so the table matches it to the last bit."""
    import torch

    n = NUM_TRAIN_TIMESTEPS

    def alpha_bar(t: float) -> float:
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = torch.tensor([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), 0.999) for i in range(n)],
                         dtype=torch.float32)
    abar_sqrt = torch.cumprod(1.0 - betas, dim=0).sqrt()
    a0, aT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
    abar_sqrt = (abar_sqrt - aT) * (a0 / (a0 - aT))
    abar = abar_sqrt ** 2
    alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
    betas = 1 - alphas  # diffusers stores betas and rebuilds alphas = 1 - betas
    return torch.cumprod(1.0 - betas, dim=0).numpy()


class DDIMSchedule:
    """timesteps + (alpha_t, alpha_prev) pairs for sampling; (alpha_src, alpha_dst) for inversion."""

    init_noise_sigma = 1.0

    def __init__(self, num_inference_steps: int, inverse: bool = False):
        if NUM_TRAIN_TIMESTEPS % num_inference_steps != 0 and num_inference_steps > NUM_TRAIN_TIMESTEPS:
            raise ValueError(f"num_inference_steps={num_inference_steps} out of range")
        self.alphas_cumprod = _alphas_cumprod()
        self.num_inference_steps = num_inference_steps
        self.step_ratio = NUM_TRAIN_TIMESTEPS // num_inference_steps
        grid = np.arange(num_inference_steps, dtype=np.int64) * self.step_ratio + 1
        self.inverse = inverse
        self.timesteps: List[int] = [int(t) for t in (grid if inverse else grid[::-1])]

    def alpha(self, t: int) -> float:
        """alpha-bar at level t; 1.0 below level 0 (set_alpha_to_one)."""
        return float(self.alphas_cumprod[t]) if t >= 0 else 1.0

    def step_alphas(self, t: int) -> Tuple[float, float]:
        """(alpha at the sample's level, alpha at the level the update lands on)."""
        if self.inverse:  # sample is at level t - ratio, lands on t
            return self.alpha(min(t - self.step_ratio, NUM_TRAIN_TIMESTEPS - 1)), self.alpha(t)
        return self.alpha(t), self.alpha(t - self.step_ratio)
