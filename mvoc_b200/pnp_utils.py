"""MVOC's hook layer (i2vgen-xl/pnp_utils.py) over the B200 kernels — the drop-in boundary.

Same entry points, argument meaning and out-of-band state as the reference:
  register_time_all(model, t, mask)                         pnp_utils.py:48-166
  modify_diffuser_attention_forward(unet)                   pnp_utils.py:169-560
  register_spatial_attention_pnp(model, schedule, inject_background)   :563-715
  register_temp_attention_pnp(model, schedule, inject_background)      :718-897
  register_resnet_injection(model, schedule)                :900-1037
  register_temp_conv_injection(model, schedule)             :1040-1105
  register_out_conv_injection(model, schedule)              :1108-1159
Processors keep the diffusers AttnProcessor call signature (:565-575) and the attributes
``injection_schedule``, ``inject_background``, ``t``, ``mask``.

Differences that are deliberate: ``// 5`` is generalised to ``n_obj + 3`` branches
(D1; identical for two objects); binary-mask blends are exact selects on u8 masks and the float-mask
blend is one fp32 lerp rounded once (D2); module indices are derived from the UNet layout instead of
being hard-coded, so reduced UNets work.  The blend + attention run as sm_100a kernels
(mvoc_qk_blend, mvoc_attn_fwd / mvoc_attn_temporal_fwd, mvoc_feature_blend); nothing here falls back
to torch math.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .unet3d import AttnProcessor2_0, _run_attention

MaskPair = Tuple[torch.Tensor, torch.Tensor]


# --------------------------------------------------------------------------
# mask preparation (once per resolution, cached): token-ordered masks for the kernels
# --------------------------------------------------------------------------
class _MaskCache:
    """The reference re-materialises a [T,h,w,C] mask per object per layer per step
    (pnp_utils.py:648-656, :805-809).  The kernels read one value per token, so the nearest-resized,
    token-ordered masks are built once per (mask list, resolution, mode) and reused."""

    def __init__(self):
        self._store = {}

    @staticmethod
    def _key(mask: Sequence[MaskPair], kind: str, h: int, w: int):
        return (kind, h, w) + tuple((m[0].data_ptr(), m[1].data_ptr(), m[0]._version, m[1]._version) for m in mask)

    def spatial_tokens(self, mask, h, w) -> torch.Tensor:
        """[n_obj, T*h*w] uint8, (frame, pixel) order, from the BINARY masks (pnp_utils.py:648-651)."""
        key = self._key(mask, "spa", h, w)
        out = self._store.get(key)
        if out is None:
            rows = []
            for _, mb in mask:
                m = mb[0].to(torch.float32)                       # [4,T,H,W]  ("a b l h w -> (a b) l h w")
                m = F.interpolate(m, size=(h, w), mode="nearest")[0]  # channel 0: [T,h,w]
                rows.append((m != 0).to(torch.uint8).reshape(-1))
            out = torch.stack(rows).contiguous()
            self._store[key] = out
        return out

    def temporal_tokens(self, mask, h, w) -> torch.Tensor:
        """[n_obj, h*w*T] float32, (pixel, frame) order, from the FLOAT masks (pnp_utils.py:805-809)."""
        key = self._key(mask, "tmp", h, w)
        out = self._store.get(key)
        if out is None:
            rows = []
            for mf, _ in mask:
                m = mf[0].to(torch.float32)                       # squeeze(0): [4,T,H,W]
                m = F.interpolate(m, size=(h, w), mode="nearest")[0]  # [T,h,w]
                rows.append(m.permute(1, 2, 0).reshape(-1))
            out = torch.stack(rows).contiguous()
            self._store[key] = out
        return out

    def feature_planes(self, mask) -> torch.Tensor:
        """[n_obj, T, H*W] uint8 from the BINARY masks at full latent resolution (pnp_utils.py:986-994)."""
        key = self._key(mask, "feat", 0, 0)
        out = self._store.get(key)
        if out is None:
            out = torch.stack([mb[0, 0].reshape(mb.shape[2], -1).to(torch.uint8) for _, mb in mask]).contiguous()
            self._store[key] = out
        return out


_MASKS = _MaskCache()


def _fires(obj) -> bool:
    """`self.t in self.injection_schedule or self.t == 1000` (pnp_utils.py:624, :778, :970, :1059, :1114)
    without a device sync: the schedule is turned into a Python set once."""
    sched = getattr(obj, "injection_schedule", None)
    if sched is None:
        return False
    cached = getattr(obj, "_schedule_cache", None)
    if cached is None or cached[0] is not sched:
        vals = sched.tolist() if isinstance(sched, torch.Tensor) else list(sched)
        cached = (sched, frozenset(int(v) for v in vals))
        obj._schedule_cache = cached
    t = int(obj.t)
    return t in cached[1] or t == 1000


# --------------------------------------------------------------------------
# processors
# --------------------------------------------------------------------------
class _InjectingProcessor(AttnProcessor2_0):
    temporal = False

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 height=None, width=None, scale: float = 1.0):
        if encoder_hidden_states is not None or not _fires(self):
            return super().__call__(attn, hidden_states, encoder_hidden_states, attention_mask, temb, scale)
        mask = self.mask
        n_obj = len(mask)
        nb = n_obj + 3
        if hidden_states.shape[0] % nb != 0:
            raise ValueError(f"batch {hidden_states.shape[0]} is not divisible into n_obj+3={nb} branches")
        # separate projections: the blend kernel wants slot-major contiguous Q and K
        q = attn.to_q(hidden_states)                                            # :604
        k = attn.to_k(hidden_states)                                            # :611
        v = attn.to_v(hidden_states)                                            # :612
        if self.temporal:
            tokens = _MASKS.temporal_tokens(mask, height, width)
        else:
            tokens = _MASKS.spatial_tokens(mask, height, width)
        ops.qk_blend_(q, k, tokens, n_obj, bool(self.inject_background))         # :628-672 / :782-850
        out = _run_attention(q, k, v, attn.heads)                               # :684 / :862
        return attn.to_out[0](out)                                              # :692


def register_spatial_attention_pnp(model, injection_schedule, inject_background=False):
    class ModifiedSpaAttnProcessor(_InjectingProcessor):
        temporal = False

    for bi, li in injected_attention_sites(model.unet):
        module = model.unet.up_blocks[bi].attentions[li].transformer_blocks[0].attn1
        p = ModifiedSpaAttnProcessor()
        setattr(p, "injection_schedule", injection_schedule)
        setattr(p, "inject_background", inject_background)
        module.processor = p


def register_temp_attention_pnp(model, injection_schedule, inject_background=False):
    class ModifiedTmpAttnProcessor(_InjectingProcessor):
        temporal = True

    for bi, li in injected_attention_sites(model.unet):
        module = model.unet.up_blocks[bi].temp_attentions[li].transformer_blocks[0].attn1
        p = ModifiedTmpAttnProcessor()
        setattr(p, "injection_schedule", injection_schedule)
        setattr(p, "inject_background", inject_background)
        module.processor = p


def injected_attention_sites(unet):
    """res_dict = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]} (pnp_utils.py:706, :889) expressed on the layout:
    every cross-attention up block, skipping layer 0 of the lowest-resolution one (comment at :707)."""
    sites, first = [], True
    for bi, blk in enumerate(unet.up_blocks):
        if not getattr(blk, "has_cross_attention", False):
            continue
        for li in range(len(blk.attentions)):
            if first and li == 0:
                continue
            sites.append((bi, li))
        first = False
    return sites


# --------------------------------------------------------------------------
# feature (conv) injection
# --------------------------------------------------------------------------
def _feature_hook(module, hidden_states):
    """Blend of the composite slots from the background slot + objects, in place
    (pnp_utils.py:970-1004 / :1059-1082 / :1114-1146)."""
    if not _fires(module):
        return
    mask = module.mask
    n_obj = len(mask)
    nb = n_obj + 3
    if hidden_states.shape[0] % nb != 0:
        raise ValueError(f"batch {hidden_states.shape[0]} is not divisible into n_obj+3={nb} branches")
    frames = hidden_states.shape[0] // nb
    ops.feature_blend_(hidden_states, _MASKS.feature_planes(mask), n_obj, frames)


def register_resnet_injection(model, injection_schedule):
    blk = model.unet.up_blocks[-1]          # up_blocks_id = [3]  (pnp_utils.py:1031)
    for m in blk.resnets:
        m.feature_hook = _feature_hook
        setattr(m, "injection_schedule", injection_schedule)


def register_temp_conv_injection(model, injection_schedule):
    blk = model.unet.up_blocks[-1]          # pnp_utils.py:1099
    for m in blk.temp_convs:
        m.feature_hook = _feature_hook
        setattr(m, "injection_schedule", injection_schedule)


def register_out_conv_injection(model, injection_schedule):
    m = model.unet.conv_out                 # pnp_utils.py:1157
    m.feature_hook = _feature_hook
    setattr(m, "injection_schedule", injection_schedule)


def modify_diffuser_attention_forward(unet):
    """The reference re-points the forward of every TransformerTemporalModel / BasicTransformerBlock /
    Attention / Transformer2DModel so that height/width reach the processors (pnp_utils.py:550-560).
    The modules of mvoc_b200.unet3d already implement exactly those forwards; this validates the tree
    instead of patching it."""
    from .unet3d import Attention, BasicTransformerBlock, Transformer2DModel, TransformerTemporalModel

    n = 0
    for _, module in unet.named_modules():
        if isinstance(module, (Attention, BasicTransformerBlock, Transformer2DModel, TransformerTemporalModel)):
            n += 1
    if n == 0:
        raise TypeError("modify_diffuser_attention_forward: not an mvoc_b200.unet3d.I2VGenXLUNet")
    return unet


# --------------------------------------------------------------------------
# per-step state broadcast
# --------------------------------------------------------------------------
def register_time_all(model, t, mask):
    """setattr(module, 't'/'mask') on every hook site (pnp_utils.py:48-166), walking the layout."""
    unet = model.unet
    t = int(t)
    for blk in unet.up_blocks:                                   # :50-62
        for m in list(blk.resnets) + list(blk.temp_convs):
            setattr(m, "t", t)
            setattr(m, "mask", mask)
    for blk in list(unet.down_blocks) + [unet.mid_block] + list(unet.up_blocks):   # :64-156
        if not getattr(blk, "has_cross_attention", False):
            continue
        for tr in list(blk.attentions) + list(blk.temp_attentions):
            tb = tr.transformer_blocks[0]
            for a in (tb.attn1, tb.attn2):
                setattr(a.processor, "t", t)
                setattr(a.processor, "mask", mask)
    for m in (unet.conv_out, unet.conv_in):                      # :158-166
        setattr(m, "t", t)
        setattr(m, "mask", mask)


def register_time(model, t):
    """Older helper kept for API parity (pnp_utils.py:36-45)."""
    for bi, li in injected_attention_sites(model.unet):
        for grp in ("attentions", "temp_attentions"):
            p = getattr(model.unet.up_blocks[bi], grp)[li].transformer_blocks[0].attn1.processor
            setattr(p, "t", int(t))
