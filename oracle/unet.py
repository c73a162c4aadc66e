"""ORACLE — test infrastructure only (never imported by the product path).

fp32 PyTorch restatement of the ``diffusers==0.27.2`` modules the reference builds its UNet
from (``I2VGenXLUNet`` and its blocks).  diffusers is an un-vendored dependency of the
reference (environment.yaml:58) and is absent here, so module constructors, block order and
the stock processor are restated from the published 0.27.2 sources —
**parity for these internals is unpinned by the reference**; they are pinned structurally
(parameter count 1 420 469 224 for the i2vgen-xl config, 145 context tokens, state-dict names).

The forward bodies that MVOC re-points (``pnp_utils.py:170-548``) and the UNet driver
(``pipelines/pipeline_i2vgen_xl.py:109-362``) ARE in the reference tree; the classes below use
attribute names identical to diffusers' so that ``tests/golden/make_golden.py`` can run the
reference's own functions over this tree (with the ``diffusers`` import stubbed to these
classes) and pin ``oracle/hooks.py`` against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock3D",) * 3 + ("DownBlock3D",)
    up_block_types: Tuple[str, ...] = ("UpBlock3D",) + ("CrossAttnUpBlock3D",) * 3
    layers_per_block: int = 2
    norm_num_groups: int = 32
    cross_attention_dim: int = 1024
    attention_head_dim: int = 64      # diffusers passes `num_attention_heads=64` and uses it as head_dim
    transformer_in_heads: int = 8

    @staticmethod
    def full() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def reduced() -> "UNetConfig":
        """BASELINE config 1: two levels, head_dim 64 kept (heads 1/2)."""
        return UNetConfig(
            block_out_channels=(64, 128),
            down_block_types=("CrossAttnDownBlock3D", "DownBlock3D"),
            up_block_types=("UpBlock3D", "CrossAttnUpBlock3D"),
            transformer_in_heads=2,
        )

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


# ------------------------------------------------------------------ embeddings
class Timesteps(nn.Module):
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0)."""

    def __init__(self, num_channels: int):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - 0.0)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip: [cos | sin]


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# ------------------------------------------------------------------ attention
class AttnProcessor2_0:
    """Stock processor = pnp_utils.py:576-612 + :674-704 without the `# Modified here` block."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0):
        batch_size = hidden_states.shape[0]
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        inner_dim = key.shape[-1]
        head_dim = inner_dim // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        hidden_states = F.scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0,
                                                       is_causal=False)
        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim)
        hidden_states = hidden_states.to(query.dtype)
        hidden_states = attn.to_out[0](hidden_states)
        hidden_states = attn.to_out[1](hidden_states)
        return hidden_states / attn.rescale_output_factor


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8,
                 dim_head: int = 64, bias: bool = False, out_bias: bool = True):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.rescale_output_factor = 1.0
        self.residual_connection = False
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(kv_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(kv_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states))


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4, activation_fn: str = "geglu"):
        super().__init__()
        inner = dim * mult
        act = GEGLU(dim, inner) if activation_fn == "geglu" else GELU(dim, inner)
        self.net = nn.ModuleList([act, nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, hidden_states):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    """forward == basic_transformer_block_forward (pnp_utils.py:222-346), norm_type 'layer_norm'."""

    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int,
                 cross_attention_dim: Optional[int] = None, double_self_attention: bool = False):
        super().__init__()
        self.only_cross_attention = False
        self.norm_type = "layer_norm"
        self.pos_embed = None
        self._chunk_size = None
        self._chunk_dim = 0
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, num_attention_heads, attention_head_dim, bias=False)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, None if double_self_attention else cross_attention_dim,
                               num_attention_heads, attention_head_dim, bias=False)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim, activation_fn="geglu")

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, timestep=None, cross_attention_kwargs=None, class_labels=None,
                height=None, width=None, added_cond_kwargs=None):
        norm_hidden_states = self.norm1(hidden_states)
        kw = {}
        if _accepts_hw(self.attn1.processor):
            kw = dict(height=height, width=width)
        attn_output = self.attn1(norm_hidden_states, encoder_hidden_states=None, attention_mask=attention_mask, **kw)
        hidden_states = attn_output + hidden_states
        norm_hidden_states = self.norm2(hidden_states)
        attn_output = self.attn2(norm_hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 attention_mask=encoder_attention_mask)
        hidden_states = attn_output + hidden_states
        norm_hidden_states = self.norm3(hidden_states)
        hidden_states = self.ff(norm_hidden_states) + hidden_states
        return hidden_states


def _accepts_hw(processor) -> bool:
    """attention_forward (pnp_utils.py:364-385) forwards height/width iff the processor takes them."""
    import inspect

    return "height" in inspect.signature(processor.__call__).parameters


class Transformer2DModel(nn.Module):
    """forward == transformer2dmodel_forward live branch (pnp_utils.py:426-434, :462-508)."""

    def __init__(self, num_attention_heads: int, attention_head_dim: int, in_channels: int,
                 cross_attention_dim: int, norm_num_groups: int = 32):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.is_input_continuous, self.is_input_vectorized, self.is_input_patches = True, False, False
        self.use_linear_projection = False
        self.caption_projection = None
        self.gradient_checkpointing = False
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, added_cond_kwargs=None,
                class_labels=None, cross_attention_kwargs=None, attention_mask=None,
                encoder_attention_mask=None, return_dict: bool = True):
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        hidden_states = self.proj_in(hidden_states)
        inner_dim = hidden_states.shape[1]
        hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states=encoder_hidden_states, height=height,
                                  width=width)
        hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        hidden_states = self.proj_out(hidden_states)
        return (hidden_states + residual,)


class TransformerTemporalModel(nn.Module):
    """forward == transformer_temporal_model_forward (pnp_utils.py:170-220)."""

    def __init__(self, num_attention_heads: int, attention_head_dim: int, in_channels: int,
                 cross_attention_dim: int, norm_num_groups: int = 32):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim,
                                   double_self_attention=True)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                num_frames: int = 1, cross_attention_kwargs=None, return_dict: bool = True):
        batch_frames, channel, height, width = hidden_states.shape
        batch_size = batch_frames // num_frames
        residual = hidden_states
        hidden_states = hidden_states[None, :].reshape(batch_size, num_frames, channel, height, width)
        hidden_states = hidden_states.permute(0, 2, 1, 3, 4)
        hidden_states = self.norm(hidden_states)
        hidden_states = hidden_states.permute(0, 3, 4, 2, 1).reshape(batch_size * height * width, num_frames, channel)
        hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states=encoder_hidden_states, height=height,
                                  width=width)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = (hidden_states[None, None, :].reshape(batch_size, height, width, num_frames, channel)
                         .permute(0, 3, 4, 1, 2).contiguous())
        hidden_states = hidden_states.reshape(batch_frames, channel, height, width)
        return (hidden_states + residual,)


# ------------------------------------------------------------------ conv blocks
class ResnetBlock2D(nn.Module):
    """forward == the closure at pnp_utils.py:902-1020 without the injection block."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int = 32,
                 eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.upsample = self.downsample = None
        self.skip_time_act = False
        self.time_embedding_norm = "default"
        self.output_scale_factor = 1.0
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, input_tensor, temb, scale: float = 1.0):
        hidden_states = self.nonlinearity(self.norm1(input_tensor))
        hidden_states = self.conv1(hidden_states)
        temb = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        hidden_states = hidden_states + temb
        hidden_states = self.nonlinearity(self.norm2(hidden_states))
        hidden_states = self.conv2(self.dropout(hidden_states))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + hidden_states) / self.output_scale_factor


class TemporalConvLayer(nn.Module):
    """forward == the closure at pnp_utils.py:1042-1057 without the injection block."""

    def __init__(self, in_dim: int, out_dim: Optional[int] = None, dropout: float = 0.1, norm_num_groups: int = 32):
        super().__init__()
        out_dim = out_dim or in_dim
        self.conv1 = nn.Sequential(nn.GroupNorm(norm_num_groups, in_dim), nn.SiLU(),
                                   nn.Conv3d(in_dim, out_dim, (3, 1, 1), padding=(1, 0, 0)))
        self.conv2 = nn.Sequential(nn.GroupNorm(norm_num_groups, out_dim), nn.SiLU(), nn.Dropout(dropout),
                                   nn.Conv3d(out_dim, in_dim, (3, 1, 1), padding=(1, 0, 0)))
        self.conv3 = nn.Sequential(nn.GroupNorm(norm_num_groups, out_dim), nn.SiLU(), nn.Dropout(dropout),
                                   nn.Conv3d(out_dim, in_dim, (3, 1, 1), padding=(1, 0, 0)))
        self.conv4 = nn.Sequential(nn.GroupNorm(norm_num_groups, out_dim), nn.SiLU(), nn.Dropout(dropout),
                                   nn.Conv3d(out_dim, in_dim, (3, 1, 1), padding=(1, 0, 0)))
        nn.init.zeros_(self.conv4[-1].weight)
        nn.init.zeros_(self.conv4[-1].bias)

    def forward(self, hidden_states, num_frames: int = 1):
        hidden_states = (hidden_states[None, :].reshape((-1, num_frames) + hidden_states.shape[1:])
                         .permute(0, 2, 1, 3, 4))
        identity = hidden_states
        hidden_states = self.conv4(self.conv3(self.conv2(self.conv1(hidden_states))))
        hidden_states = identity + hidden_states
        return hidden_states.permute(0, 2, 1, 3, 4).reshape(
            (hidden_states.shape[0] * hidden_states.shape[2], -1) + hidden_states.shape[3:])


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, hidden_states, scale: float = 1.0):
        return self.conv(hidden_states)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, scale: float = 1.0):
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        return self.conv(hidden_states)


class _Block3D(nn.Module):
    has_cross_attention = False

    def _make(self, n, in_chs, out_ch, temb, heads_dim, cross_dim, groups, attn: bool):
        self.resnets = nn.ModuleList([ResnetBlock2D(ic, out_ch, temb, groups) for ic in in_chs])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(out_ch, out_ch, 0.1, groups) for _ in range(n)])
        if attn:
            heads = out_ch // heads_dim
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, heads_dim, out_ch, cross_dim, groups) for _ in range(n)])
            self.temp_attentions = nn.ModuleList(
                [TransformerTemporalModel(heads, heads_dim, out_ch, cross_dim, groups) for _ in range(n)])


class CrossAttnDownBlock3D(_Block3D):
    has_cross_attention = True

    def __init__(self, in_ch, out_ch, temb, n_layers, head_dim, cross_dim, groups, add_downsample):
        super().__init__()
        self._make(n_layers, [in_ch] + [out_ch] * (n_layers - 1), out_ch, temb, head_dim, cross_dim, groups, True)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                num_frames: int = 1, cross_attention_kwargs=None):
        output_states = ()
        for resnet, temp_conv, attn, temp_attn in zip(self.resnets, self.temp_convs, self.attentions,
                                                      self.temp_attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            hidden_states = temp_attn(hidden_states, num_frames=num_frames,
                                      cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            output_states += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class DownBlock3D(_Block3D):
    def __init__(self, in_ch, out_ch, temb, n_layers, groups, add_downsample):
        super().__init__()
        self._make(n_layers, [in_ch] + [out_ch] * (n_layers - 1), out_ch, temb, 0, 0, groups, False)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, num_frames: int = 1):
        output_states = ()
        for resnet, temp_conv in zip(self.resnets, self.temp_convs):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
            output_states += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class UNetMidBlock3DCrossAttn(_Block3D):
    has_cross_attention = True

    def __init__(self, ch, temb, head_dim, cross_dim, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups) for _ in range(2)])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(ch, ch, 0.1, groups) for _ in range(2)])
        heads = ch // head_dim
        self.attentions = nn.ModuleList([Transformer2DModel(heads, head_dim, ch, cross_dim, groups)])
        self.temp_attentions = nn.ModuleList([TransformerTemporalModel(heads, head_dim, ch, cross_dim, groups)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                num_frames: int = 1, cross_attention_kwargs=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.temp_convs[0](hidden_states, num_frames=num_frames)
        for attn, temp_attn, resnet, temp_conv in zip(self.attentions, self.temp_attentions, self.resnets[1:],
                                                      self.temp_convs[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            hidden_states = temp_attn(hidden_states, num_frames=num_frames,
                                      cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
        return hidden_states


class CrossAttnUpBlock3D(_Block3D):
    has_cross_attention = True

    def __init__(self, in_ch, out_ch, prev_out, temb, n_layers, head_dim, cross_dim, groups, add_upsample):
        super().__init__()
        ins = []
        for i in range(n_layers):
            res_skip = in_ch if i == n_layers - 1 else out_ch
            res_in = prev_out if i == 0 else out_ch
            ins.append(res_in + res_skip)
        self._make(n_layers, ins, out_ch, temb, head_dim, cross_dim, groups, True)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                upsample_size=None, attention_mask=None, num_frames: int = 1, cross_attention_kwargs=None):
        for resnet, temp_conv, attn, temp_attn in zip(self.resnets, self.temp_convs, self.attentions,
                                                      self.temp_attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
            hidden_states = temp_attn(hidden_states, num_frames=num_frames,
                                      cross_attention_kwargs=cross_attention_kwargs, return_dict=False)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class UpBlock3D(_Block3D):
    def __init__(self, in_ch, out_ch, prev_out, temb, n_layers, groups, add_upsample):
        super().__init__()
        ins = []
        for i in range(n_layers):
            res_skip = in_ch if i == n_layers - 1 else out_ch
            res_in = prev_out if i == 0 else out_ch
            ins.append(res_in + res_skip)
        self._make(n_layers, ins, out_ch, temb, 0, 0, groups, False)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, num_frames: int = 1):
        for resnet, temp_conv in zip(self.resnets, self.temp_convs):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class I2VGenXLTransformerTemporalEncoder(nn.Module):
    def __init__(self, dim: int, num_attention_heads: int = 2, attention_head_dim: int = 4):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, num_attention_heads, attention_head_dim, bias=False, out_bias=True)
        self.ff = FeedForward(dim, activation_fn="gelu")

    def forward(self, hidden_states):
        norm_hidden_states = self.norm1(hidden_states)
        attn_output = self.attn1(norm_hidden_states, encoder_hidden_states=None)
        hidden_states = attn_output + hidden_states
        if hidden_states.ndim == 4:
            hidden_states = hidden_states.squeeze(1)
        return self.ff(hidden_states) + hidden_states


class _Cfg:
    def __init__(self, c: UNetConfig):
        self.in_channels = c.in_channels
        self.cross_attention_dim = c.cross_attention_dim


class I2VGenXLUNet(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        self.config = _Cfg(cfg)
        boc = cfg.block_out_channels
        temb = cfg.time_embed_dim
        g, hd, cd = cfg.norm_num_groups, cfg.attention_head_dim, cfg.cross_attention_dim
        ic = cfg.in_channels
        self.conv_in = nn.Conv2d(ic + ic, boc[0], 3, padding=1)
        self.transformer_in = TransformerTemporalModel(cfg.transformer_in_heads, hd, boc[0], cd, g)
        self.image_latents_proj_in = nn.Sequential(
            nn.Conv2d(4, ic * 4, 3, padding=1), nn.SiLU(),
            nn.Conv2d(ic * 4, ic * 4, 3, stride=1, padding=1), nn.SiLU(),
            nn.Conv2d(ic * 4, ic, 3, stride=1, padding=1))
        self.image_latents_temporal_encoder = I2VGenXLTransformerTemporalEncoder(ic, 2, ic)
        self.image_latents_context_embedding = nn.Sequential(
            nn.Conv2d(4, ic * 8, 3, padding=1), nn.SiLU(), nn.AdaptiveAvgPool2d((32, 32)),
            nn.Conv2d(ic * 8, ic * 16, 3, stride=2, padding=1), nn.SiLU(),
            nn.Conv2d(ic * 16, cd, 3, stride=2, padding=1))
        self.time_proj = Timesteps(boc[0])
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.context_embedding = nn.Sequential(nn.Linear(cd, temb), nn.SiLU(), nn.Linear(temb, cd * ic))
        self.fps_embedding = nn.Sequential(nn.Linear(boc[0], temb), nn.SiLU(), nn.Linear(temb, temb))

        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            final = i == len(boc) - 1
            if t == "CrossAttnDownBlock3D":
                self.down_blocks.append(CrossAttnDownBlock3D(in_ch, out_ch, temb, cfg.layers_per_block, hd, cd, g,
                                                             not final))
            else:
                self.down_blocks.append(DownBlock3D(in_ch, out_ch, temb, cfg.layers_per_block, g, not final))
        self.mid_block = UNetMidBlock3DCrossAttn(boc[-1], temb, hd, cd, g)
        self.num_upsamplers = 0
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, t in enumerate(cfg.up_block_types):
            final = i == len(boc) - 1
            prev_out, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            add_up = not final
            if add_up:
                self.num_upsamplers += 1
            if t == "CrossAttnUpBlock3D":
                self.up_blocks.append(CrossAttnUpBlock3D(in_ch, out_ch, prev_out, temb, cfg.layers_per_block + 1,
                                                         hd, cd, g, add_up))
            else:
                self.up_blocks.append(UpBlock3D(in_ch, out_ch, prev_out, temb, cfg.layers_per_block + 1, g, add_up))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, fps, image_latents, image_embeddings=None, encoder_hidden_states=None,
                timestep_cond=None, cross_attention_kwargs=None, return_dict: bool = True):
        """Stock diffusers forward, used by invert()/__call__ (pipeline:1173, :1952)."""
        batch_size, channels, num_frames, height, width = sample.shape
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % up_factor != 0 for s in sample.shape[-2:])
        upsample_size = None
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_embedding(self.time_proj(timesteps).to(self.dtype), timestep_cond)
        fps = fps.expand(fps.shape[0])
        fps_emb = self.fps_embedding(self.time_proj(fps).to(self.dtype))
        emb = (t_emb + fps_emb).repeat_interleave(repeats=num_frames, dim=0)

        context_emb = sample.new_zeros(batch_size, 0, self.config.cross_attention_dim)
        context_emb = torch.cat([context_emb, encoder_hidden_states], dim=1)
        il0 = image_latents[:, :, :1, :]
        il0 = il0.permute(0, 2, 1, 3, 4).reshape(il0.shape[0] * il0.shape[2], il0.shape[1], il0.shape[3], il0.shape[4])
        il0 = self.image_latents_context_embedding(il0)
        _b, _c, _h, _w = il0.shape
        il0 = il0.permute(0, 2, 3, 1).reshape(_b, _h * _w, _c)
        context_emb = torch.cat([context_emb, il0], dim=1)
        image_emb = self.context_embedding(image_embeddings)
        image_emb = image_emb.view(-1, self.config.in_channels, self.config.cross_attention_dim)
        context_emb = torch.cat([context_emb, image_emb], dim=1)
        context_emb = context_emb.repeat_interleave(repeats=num_frames, dim=0)

        il = image_latents.permute(0, 2, 1, 3, 4).reshape(batch_size * num_frames, channels, height, width)
        il = self.image_latents_proj_in(il)
        il = (il[None, :].reshape(batch_size, num_frames, channels, height, width).permute(0, 3, 4, 1, 2)
              .reshape(batch_size * height * width, num_frames, channels))
        il = self.image_latents_temporal_encoder(il)
        il = il.reshape(batch_size, height, width, num_frames, channels).permute(0, 4, 3, 1, 2)

        sample = torch.cat([sample, il], dim=1)
        sample = sample.permute(0, 2, 1, 3, 4).reshape((sample.shape[0] * num_frames, -1) + sample.shape[3:])
        sample = self.conv_in(sample)
        sample = self.transformer_in(sample, num_frames=num_frames, return_dict=False)[0]
        return unet_body(self, sample, emb, context_emb, num_frames, forward_upsample_size, return_dict)


def unet_body(model, sample, emb, context_emb, num_frames, forward_upsample_size, return_dict=False):
    """Down / mid / up / out traversal, pipelines/pipeline_i2vgen_xl.py:292-362."""
    upsample_size = None
    down_block_res_samples = (sample,)
    for blk in model.down_blocks:
        if getattr(blk, "has_cross_attention", False):
            sample, res = blk(hidden_states=sample, temb=emb, encoder_hidden_states=context_emb,
                              num_frames=num_frames)
        else:
            sample, res = blk(hidden_states=sample, temb=emb, num_frames=num_frames)
        down_block_res_samples += res
    sample = model.mid_block(sample, emb, encoder_hidden_states=context_emb, num_frames=num_frames)
    for i, blk in enumerate(model.up_blocks):
        is_final = i == len(model.up_blocks) - 1
        n = len(blk.resnets)
        res = down_block_res_samples[-n:]
        down_block_res_samples = down_block_res_samples[:-n]
        if not is_final and forward_upsample_size:
            upsample_size = down_block_res_samples[-1].shape[2:]
        if getattr(blk, "has_cross_attention", False):
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                         encoder_hidden_states=context_emb, upsample_size=upsample_size, num_frames=num_frames)
        else:
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                         upsample_size=upsample_size, num_frames=num_frames)
    sample = model.conv_act(model.conv_norm_out(sample))
    sample = model.conv_out(sample)
    sample = sample[None, :].reshape((-1, num_frames) + sample.shape[1:]).permute(0, 2, 1, 3, 4)
    return (sample,)


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())
