"""ORACLE — test infrastructure only (never imported by mvoc_b200/).

Decoder side of diffusers 0.27.2 ``AutoencoderKL`` (the stabilityai/sd-vae architecture i2vgen-xl ships) restated
in fp32, and the reference's ``decode_latents`` (pipelines/pipeline_i2vgen_xl.py:771-791) on top of it.  It
exists for ONE purpose: SURVEY §8(d) expresses end-to-end parity also as the PSNR of frames decoded "by the same
(oracle-side, torch) decoder from both latents".  There is no checkpoint in this environment, so the decoder is
random-init (seeded) — the PSNR it yields measures how far two latents are apart through a VAE-shaped map, it is
not an image-quality number.  diffusers is un-vendored: the module layout follows the published architecture
(parity unpinned by the reference, see oracle/unet.py).

Layout: post_quant_conv(4->4, 1x1) -> conv_in(4->C3) -> mid (resnet, 1-head attention, resnet) ->
4 up blocks of 3 resnets (+ nearest-x2 upsample conv, except the last) with channels C3, C3, C2, C1 -> GroupNorm(32)
-> SiLU -> conv_out(C0->3).  The real model has (C0..C3) = (128, 256, 512, 512); tests use a narrow one.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

SCALING_FACTOR = 0.18215


class _Resnet(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return h + (x if self.conv_shortcut is None else self.conv_shortcut(x))


class _MidAttention(nn.Module):
    """Single-head self-attention over the pixels of the lowest-resolution map (diffusers Attention with
    residual_connection=True, GroupNorm input)."""

    def __init__(self, c: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(c), dim=-1) @ v
        return x + self.to_out[0](a).transpose(1, 2).reshape(b, c, h, w)


class _UpBlock(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int, upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(3)])
        self.upsample = nn.Conv2d(cout, cout, 3, padding=1) if upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsample is not None:
            x = self.upsample(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class DecoderKL(nn.Module):
    def __init__(self, block_out_channels: Sequence[int] = (128, 256, 512, 512), groups: int = 32):
        super().__init__()
        c = list(block_out_channels)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.conv_in = nn.Conv2d(4, c[-1], 3, padding=1)
        self.mid_resnet1 = _Resnet(c[-1], c[-1], groups)
        self.mid_attn = _MidAttention(c[-1], groups)
        self.mid_resnet2 = _Resnet(c[-1], c[-1], groups)
        rev = c[::-1]
        self.up_blocks = nn.ModuleList(
            [_UpBlock(rev[max(i - 1, 0)], rev[i], groups, upsample=i < len(rev) - 1) for i in range(len(rev))])
        self.conv_norm_out = nn.GroupNorm(groups, c[0], eps=1e-6)
        self.conv_out = nn.Conv2d(c[0], 3, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(self.post_quant_conv(z))
        x = self.mid_resnet2(self.mid_attn(self.mid_resnet1(x)))
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def build_decoder(narrow: bool = True, seed: int = 0) -> DecoderKL:
    torch.manual_seed(seed)
    m = DecoderKL((32, 64, 128, 128), groups=8) if narrow else DecoderKL()
    return m.eval().requires_grad_(False)


@torch.no_grad()
def decode_latents(decoder: DecoderKL, latents: torch.Tensor, decode_chunk_size: int = 4) -> torch.Tensor:
    """pipelines/pipeline_i2vgen_xl.py:771-791: [b, 4, T, h, w] latents -> [b, 3, T, 8h, 8w] video, fp32."""
    latents = 1 / SCALING_FACTOR * latents.float()
    b, c, f, h, w = latents.shape
    flat = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    frames = [decoder(flat[i:i + decode_chunk_size]) for i in range(0, flat.shape[0], decode_chunk_size)]
    image = torch.cat(frames, dim=0)
    return image[None, :].reshape((b, f, -1) + image.shape[2:]).permute(0, 2, 1, 3, 4).float()


def psnr(video: torch.Tensor, reference: torch.Tensor) -> float:
    """PSNR in dB with the reference's dynamic range as the peak (the decoder is random-init: there is no fixed
    [-1, 1] image range to lean on)."""
    mse = float(((video.double() - reference.double()) ** 2).mean())
    peak = float(reference.max() - reference.min())
    return float("inf") if mse == 0.0 else 10.0 * math.log10(peak * peak / mse)
