"""ORACLE — test infrastructure only.

Restatement of diffusers 0.27.2 ``DDIMScheduler`` / ``DDIMInverseScheduler`` with the i2vgen-xl
``scheduler_config.json``: 1000 train steps, ``squaredcos_cap_v2`` betas,
``rescale_betas_zero_snr``, ``v_prediction``, ``timestep_spacing='leading'``, ``steps_offset=1``,
``clip_sample=False``, ``set_alpha_to_one=True``, eta = 0.  diffusers is un-vendored, so these follow
the published algorithm (parity unpinned by the reference); pinned by the known answers of
SURVEY App. A.6 in tests/test_scheduler.py.  Called from pipelines/pipeline_i2vgen_xl.py:1552-1555,
:1728 (composition) and :1914-1915, :1979 (inversion).
"""
from __future__ import annotations

import math

import torch

NUM_TRAIN = 1000


def _betas_cosine(n: int = NUM_TRAIN, max_beta: float = 0.999) -> torch.Tensor:
    def alpha_bar(t: float) -> float:
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(n):
        t1, t2 = i / n, (i + 1) / n
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


def _rescale_zero_terminal_snr(betas: torch.Tensor) -> torch.Tensor:
    alphas = 1.0 - betas
    alphas_bar_sqrt = torch.cumprod(alphas, dim=0).sqrt()
    a0 = alphas_bar_sqrt[0].clone()
    aT = alphas_bar_sqrt[-1].clone()
    alphas_bar_sqrt = alphas_bar_sqrt - aT
    alphas_bar_sqrt = alphas_bar_sqrt * (a0 / (a0 - aT))
    alphas_bar = alphas_bar_sqrt ** 2
    alphas = alphas_bar[1:] / alphas_bar[:-1]
    alphas = torch.cat([alphas_bar[0:1], alphas])
    return 1 - alphas


def alphas_cumprod() -> torch.Tensor:
    betas = _rescale_zero_terminal_snr(_betas_cosine())
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMScheduler:
    init_noise_sigma = 1.0

    def __init__(self):
        self.alphas_cumprod = alphas_cumprod()
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = NUM_TRAIN // n
        self.timesteps = (torch.arange(0, n) * ratio).flip(0) + 1  # leading, steps_offset = 1

    def scale_model_input(self, x, t):
        return x

    def step(self, model_output, timestep: int, sample):
        prev = timestep - NUM_TRAIN // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        direction = (1 - a_prev) ** 0.5 * eps
        return a_prev ** 0.5 * x0 + direction


class DDIMInverseScheduler:
    init_noise_sigma = 1.0

    def __init__(self):
        self.alphas_cumprod = alphas_cumprod()
        self.initial_alpha_cumprod = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = NUM_TRAIN // n
        self.timesteps = torch.arange(0, n) * ratio + 1

    def scale_model_input(self, x, t):
        return x

    def step(self, model_output, timestep: int, sample):
        prev_timestep = timestep
        timestep = min(timestep - NUM_TRAIN // self.num_inference_steps, NUM_TRAIN - 1)
        a_t = self.alphas_cumprod[timestep] if timestep >= 0 else self.initial_alpha_cumprod
        a_prev = self.alphas_cumprod[prev_timestep]
        b_t = 1 - a_t
        x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps
