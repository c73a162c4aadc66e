"""ORACLE — test infrastructure only (never imported by the product path).

fp32 restatement of MVOC's hook layer, ``i2vgen-xl/pnp_utils.py``: the composite attention
processors, the resnet / temporal-conv / conv_out feature injection and the per-step state
broadcast — generalised from the hard-coded ``// 5`` (two objects) to ``n_obj + 3`` branches
(deviation D1; identical for two objects).

Pinned by ``tests/golden/``: ``make_golden.py`` runs the reference's own ``register_*``
functions over ``oracle/unet.py`` modules and ``tests/test_oracle_golden.py`` requires this file
to reproduce those outputs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops_ref
from .unet import AttnProcessor2_0


def _fires(obj) -> bool:
    """`self.t in self.injection_schedule or self.t == 1000` (pnp_utils.py:624, :778, :970, :1059, :1114)."""
    sched = obj.injection_schedule
    if sched is None:
        return False
    t = obj.t
    if isinstance(sched, torch.Tensor):
        hit = bool((sched == t).any()) if sched.numel() else False
    else:
        hit = t in sched
    return hit or t == 1000


class ModifiedSpaAttnProcessor(AttnProcessor2_0):
    """pnp_utils.py:564-704."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 height=None, width=None, scale: float = 1.0):
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        if _fires(self):
            query, key = ops_ref.spatial_qk_inject_ref(query, key, self.mask, height, width,
                                                       self.inject_background)
        out = ops_ref.sdpa_ref(query, key, value, attn.heads)
        out = attn.to_out[0](out)
        out = attn.to_out[1](out)
        return out / attn.rescale_output_factor


class ModifiedTmpAttnProcessor(AttnProcessor2_0):
    """pnp_utils.py:719-887."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 height=None, width=None, scale: float = 1.0):
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        if _fires(self):
            query, key = ops_ref.temporal_qk_inject_ref(query, key, self.mask, height, width,
                                                        self.inject_background)
        out = ops_ref.sdpa_ref(query, key, value, attn.heads)
        out = attn.to_out[0](out)
        out = attn.to_out[1](out)
        return out / attn.rescale_output_factor


def injected_attention_sites(unet):
    """(block index, layer index) pairs that get a Modified*Processor.

    Reference: res_dict = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]} (pnp_utils.py:706, :889) — "all
    cross-attention up blocks; not the first layer of the lowest-resolution one" (comment :707).
    """
    sites = []
    first = True
    for bi, blk in enumerate(unet.up_blocks):
        if not getattr(blk, "has_cross_attention", False):
            continue
        for li in range(len(blk.attentions)):
            if first and li == 0:
                continue
            sites.append((bi, li))
        first = False
    return sites


def register_spatial_attention_pnp(model, injection_schedule, inject_background=False):
    for bi, li in injected_attention_sites(model.unet):
        module = model.unet.up_blocks[bi].attentions[li].transformer_blocks[0].attn1
        p = ModifiedSpaAttnProcessor()
        p.injection_schedule, p.inject_background = injection_schedule, inject_background
        module.processor = p


def register_temp_attention_pnp(model, injection_schedule, inject_background=False):
    for bi, li in injected_attention_sites(model.unet):
        module = model.unet.up_blocks[bi].temp_attentions[li].transformer_blocks[0].attn1
        p = ModifiedTmpAttnProcessor()
        p.injection_schedule, p.inject_background = injection_schedule, inject_background
        module.processor = p


def register_resnet_injection(model, injection_schedule):
    """pnp_utils.py:900-1037: last up block's resnets; blend after conv2, before the shortcut."""

    def conv_forward(self):
        def forward(input_tensor, temb, scale: float = 1.0):
            hidden_states = self.nonlinearity(self.norm1(input_tensor))
            hidden_states = self.conv1(hidden_states)
            t = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
            hidden_states = hidden_states + t
            hidden_states = self.nonlinearity(self.norm2(hidden_states))
            hidden_states = self.conv2(self.dropout(hidden_states))
            if _fires(self):
                hidden_states = ops_ref.feature_inject_ref(hidden_states, self.mask)
            if self.conv_shortcut is not None:
                input_tensor = self.conv_shortcut(input_tensor)
            return (input_tensor + hidden_states) / self.output_scale_factor

        return forward

    blk = model.unet.up_blocks[-1]
    for m in blk.resnets:
        m.forward = conv_forward(m)
        m.injection_schedule = injection_schedule


def register_temp_conv_injection(model, injection_schedule):
    """pnp_utils.py:1040-1105."""

    def conv_forward(self):
        def forward(hidden_states, num_frames: int = 1):
            hidden_states = (hidden_states[None, :].reshape((-1, num_frames) + hidden_states.shape[1:])
                             .permute(0, 2, 1, 3, 4))
            identity = hidden_states
            hidden_states = self.conv4(self.conv3(self.conv2(self.conv1(hidden_states))))
            hidden_states = identity + hidden_states
            hidden_states = hidden_states.permute(0, 2, 1, 3, 4).reshape(
                (hidden_states.shape[0] * hidden_states.shape[2], -1) + hidden_states.shape[3:])
            if _fires(self):
                hidden_states = ops_ref.feature_inject_ref(hidden_states, self.mask)
            return hidden_states

        return forward

    blk = model.unet.up_blocks[-1]
    for m in blk.temp_convs:
        m.forward = conv_forward(m)
        m.injection_schedule = injection_schedule


def register_out_conv_injection(model, injection_schedule):
    """pnp_utils.py:1108-1159."""

    def conv_forward(self):
        def forward(input):
            sample = self._conv_forward(input, self.weight, self.bias)
            if _fires(self):
                sample = ops_ref.feature_inject_ref(sample, self.mask)
            return sample

        return forward

    m = model.unet.conv_out
    m.forward = conv_forward(m)
    m.injection_schedule = injection_schedule


def register_time_all(model, t, mask):
    """pnp_utils.py:48-166, driven by the module layout instead of hard-coded indices."""
    unet = model.unet
    for blk in unet.up_blocks:
        for m in list(blk.resnets) + list(blk.temp_convs):
            m.t, m.mask = t, mask
    blocks = list(unet.down_blocks) + [unet.mid_block] + list(unet.up_blocks)
    for blk in blocks:
        if not getattr(blk, "has_cross_attention", False):
            continue
        for tr in list(blk.attentions) + list(blk.temp_attentions):
            tb = tr.transformer_blocks[0]
            for a in (tb.attn1, tb.attn2):
                a.processor.t, a.processor.mask = t, mask
    for m in (unet.conv_out, unet.conv_in):
        m.t, m.mask = t, mask
