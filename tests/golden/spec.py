"""Shared definition of the golden cases (inputs are regenerated from seeds; only outputs are stored)."""
from __future__ import annotations

import torch

from mvoc_b200 import synthetic
from oracle.unet import I2VGenXLUNet, UNetConfig

T, H, W, N_OBJ = 4, 16, 16, 2

CASES = [
    dict(name="all_hooks_t981", t=981, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0, inject_background=False),
    dict(name="attn_only_t481", t=481, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0, inject_background=False),
    dict(name="inject_bg_t481", t=481, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0, inject_background=True),
    dict(name="no_hooks_t21", t=21, pnp_f_t=0.2, pnp_spatial_attn_t=0.2, pnp_temp_attn_t=0.5, inject_background=False),
]


def tiny4_config() -> UNetConfig:
    """Full 4-level I2VGenXL layout (what the reference's hard-coded indices need), narrow channels."""
    return UNetConfig(block_out_channels=(64, 128, 256, 256), transformer_in_heads=2)


def build_tiny4(seed: int = 0) -> I2VGenXLUNet:
    torch.manual_seed(seed)
    return I2VGenXLUNet(tiny4_config()).eval().requires_grad_(False)


def timesteps_50():
    return [981 - 20 * i for i in range(50)]


def make_inputs(case) -> dict:
    wl = synthetic.Workload("tiny4", "tiny4", T, H, W, N_OBJ)
    base = synthetic.make_inputs(wl, [case["t"]], {case["t"]: 0.5})
    g = torch.Generator().manual_seed(77)
    base["sample"] = torch.randn(N_OBJ + 3, 4, T, H, W, generator=g)
    return base
