"""ORACLE — test infrastructure only (never imported by the product path).

fp32 PyTorch restatements of the per-op arithmetic on MVOC's composition hot
path.  Every function cites the reference lines it follows (paths relative to
the SobeyMIL/MVOC tree).  The MVOC hook functions here are pinned against the
reference's own code: ``tests/golden/make_golden.py`` executes the real
``pnp_utils`` processors/closures (with the absent ``diffusers`` import stubbed)
and stores their outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
replays them through this file.  Pieces that live in un-vendored
``diffusers==0.27.2`` (GroupNorm modules, schedulers) are restated from the
published algorithm — parity for those is *unpinned* by the reference.

Generalisation D1: the reference hard-codes ``batch // 5``
(two objects); here ``n_branches = n_obj + 3``; identical for ``n_obj == 2``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F

MaskPair = Tuple[torch.Tensor, torch.Tensor]  # (float [1,4,T,H,W], bool [1,4,T,H,W])


# --------------------------------------------------------------------------
# attention core
# --------------------------------------------------------------------------
def sdpa_ref(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """pnp_utils.py:674-688 — view to [B, heads, N, 64], F.scaled_dot_product_attention, merge heads."""
    B, _, inner = q.shape
    hd = inner // heads
    qh = q.view(B, -1, heads, hd).transpose(1, 2)
    kh = k.view(B, -1, heads, hd).transpose(1, 2)
    vh = v.view(B, -1, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=None, dropout_p=0.0, is_causal=False)
    return o.transpose(1, 2).reshape(B, -1, heads * hd)


# --------------------------------------------------------------------------
# MVOC injections
# --------------------------------------------------------------------------
def nearest_mask(mask_thw: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """F.interpolate(mode='nearest') as used at pnp_utils.py:650 and :807 ([T,H,W] -> [T,h,w])."""
    return F.interpolate(mask_thw[None].float(), size=(height, width), mode="nearest")[0]


def spatial_qk_inject_ref(
    query: torch.Tensor,
    key: torch.Tensor,
    masks: Sequence[MaskPair],
    height: int,
    width: int,
    inject_background: bool,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """pnp_utils.py:628-672.  query/key [n_branches*T, h*w, C]; binary mask, nearest-resized."""
    n_obj = len(masks)
    nb = n_obj + 3
    query = query.clone().view(query.shape[0], height, width, -1)
    key = key.clone().view(key.shape[0], height, width, -1)
    c = query.shape[0] // nb
    u, cn = n_obj + 1, n_obj + 2
    if inject_background:  # :633-636
        q_inject, k_inject = query[:c], key[:c]
    else:  # :637-641
        q_inject, k_inject = query[cn * c:], key[cn * c:]
    for j, (_, mask_bool) in enumerate(masks):  # :643-662
        obj_q = query[c * (j + 1): c * (j + 2)]
        obj_k = key[c * (j + 1): c * (j + 2)]
        m = mask_bool.to(torch.float32)
        m = m.reshape(-1, *m.shape[2:])  # "a b l h w -> (a b) l h w"
        m = F.interpolate(m, size=(height, width), mode="nearest")[0]  # [T,h,w]
        m = m.unsqueeze(-1)
        q_inject = q_inject * (1 - m) + obj_q * m
        k_inject = k_inject * (1 - m) + obj_k * m
    query[u * c: cn * c] = q_inject  # :664-668
    key[u * c: cn * c] = k_inject
    query[cn * c:] = q_inject
    key[cn * c:] = k_inject
    return query.view(query.shape[0], height * width, -1), key.view(key.shape[0], height * width, -1)


def temporal_qk_inject_ref(
    query: torch.Tensor,
    key: torch.Tensor,
    masks: Sequence[MaskPair],
    height: int,
    width: int,
    inject_background: bool,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """pnp_utils.py:782-850.  query/key [(n_branches h w), T, C]; float mask, nearest-resized."""
    n_obj = len(masks)
    nb = n_obj + 3
    T, C = query.shape[1], query.shape[2]
    query = query.clone().view(nb, height, width, T, C)
    key = key.clone().view(nb, height, width, T, C)
    u, cn = n_obj + 1, n_obj + 2
    if inject_background:
        q_inject, k_inject = query[:1], key[:1]
    else:
        q_inject, k_inject = query[cn:], key[cn:]
    for j, (mask_float, _) in enumerate(masks):
        obj_q, obj_k = query[j + 1: j + 2], key[j + 1: j + 2]
        m = mask_float.to(torch.float32).squeeze(0)  # [4,T,H,W]   (:805-806)
        m = F.interpolate(m, size=(height, width), mode="nearest")  # (:807)
        m = m.permute(0, 2, 3, 1)  # "b l h w -> b h w l"  (:808)
        m = m.unsqueeze(-1)[:1]  # chunk_size == 1 video   (:809)
        q_inject = q_inject * (1 - m) + obj_q * m
        k_inject = k_inject * (1 - m) + obj_k * m
    query[u: u + 1] = q_inject  # (:819-823; inside the loop in the reference, same result)
    key[u: u + 1] = k_inject
    query[cn:] = q_inject
    key[cn:] = k_inject
    return query.view(nb * height * width, T, C), key.view(nb * height * width, T, C)


def feature_inject_ref(hidden: torch.Tensor, masks: Sequence[MaskPair]) -> torch.Tensor:
    """pnp_utils.py:970-1004 / :1059-1082 / :1114-1146.  hidden [n_branches*T, C, H, W]; base = slot 0."""
    n_obj = len(masks)
    nb = n_obj + 3
    hidden = hidden.clone()
    sb = hidden.shape[0] // nb
    inject = hidden[:sb]
    for j, (_, mask_bool) in enumerate(masks):
        obj = hidden[sb * (j + 1): sb * (j + 2)]
        m = mask_bool.to(torch.float32)
        m = m.reshape(-1, *m.shape[2:])[0]  # [T,H,W]
        m = m.unsqueeze(1)
        inject = inject * (1 - m) + obj * m
    hidden[(n_obj + 1) * sb: (n_obj + 2) * sb] = inject
    hidden[(n_obj + 2) * sb:] = inject
    return hidden


# --------------------------------------------------------------------------
# normalisation
# --------------------------------------------------------------------------
def group_norm_ref(x: torch.Tensor, weight, bias, groups: int, eps: float, silu: bool,
                   frames_per_stat: int = 1) -> torch.Tensor:
    """nn.GroupNorm(+SiLU).  frames_per_stat=T restates the 5-D GroupNorm of pnp_utils.py:185-188 and
    TemporalConvLayer (:1043-1051): x is [B*T, C, H, W] and statistics span the T frames of a video."""
    if frames_per_stat == 1:
        y = F.group_norm(x, groups, weight, bias, eps)
    else:
        bt, c, h, w = x.shape
        b = bt // frames_per_stat
        x5 = x[None, :].reshape(b, frames_per_stat, c, h, w).permute(0, 2, 1, 3, 4)
        y5 = F.group_norm(x5, groups, weight, bias, eps)
        y = y5.permute(0, 2, 1, 3, 4).reshape(bt, c, h, w)
    return F.silu(y) if silu else y


# --------------------------------------------------------------------------
# latent-space step arithmetic
# --------------------------------------------------------------------------
def latent_fusion_ref(latents, bg, objs: Sequence[torch.Tensor], masks_float: Sequence[torch.Tensor],
                      ratio: float, obj_random_noise_fusion: bool = False) -> torch.Tensor:
    """pipelines/pipeline_i2vgen_xl.py:1644-1663."""
    latents = ratio * latents + (1.0 - ratio) * bg
    for obj, m in zip(objs, masks_float):
        ddim_inv_obj = obj * m
        background = latents * (1.0 - m)
        if obj_random_noise_fusion:
            foreground = latents * m
            fusion = foreground * ratio + (1 - ratio) * ddim_inv_obj
        else:
            fusion = ddim_inv_obj
        latents = background + fusion
    return latents


def ddim_step_ref(v, x, alpha_t: float, alpha_prev: float) -> torch.Tensor:
    """diffusers DDIMScheduler.step, prediction_type='v_prediction', eta=0, no clipping
    (called at pipelines/pipeline_i2vgen_xl.py:1728)."""
    a_t = torch.tensor(alpha_t, dtype=torch.float64)
    a_p = torch.tensor(alpha_prev, dtype=torch.float64)
    sa, sb = a_t.sqrt().float(), (1 - a_t).sqrt().float()
    x0 = sa * x - sb * v
    eps = sa * v + sb * x
    return a_p.sqrt().float() * x0 + (1 - a_p).sqrt().float() * eps


def cfg_ref(pred_uncond, pred_cond, guidance: float) -> torch.Tensor:
    """pipelines/pipeline_i2vgen_xl.py:1717."""
    return pred_uncond + guidance * (pred_cond - pred_uncond)
