"""I2VGenXL-architecture UNet3D, product side (bf16 on one B200).

Module/attribute names and the state-dict layout follow diffusers' ``I2VGenXLUNet`` so that MVOC's
hook functions (``mvoc_b200.pnp_utils.register_*``, mirroring i2vgen-xl/pnp_utils.py) address the same
objects — ``unet.up_blocks[i].attentions[j].transformer_blocks[0].attn1.processor`` etc. — and a real
checkpoint's state dict loads unchanged.

What runs underneath is this repo's C-ABI library (``mvoc_b200.ops``): every attention
(tcgen05 flash attention / warp-per-pixel temporal attention), every GroupNorm(+SiLU), the GEGLU
gate, the mask blends and the latent/DDIM updates are sm_100a kernels.  Dense GEMMs / convolutions
stay on cuBLAS / cuDNN through torch.  There is no CPU path: tensors must be CUDA.

Data layout: activations are channels-last ``[B*T, H, W, C]`` (frames outermost) from conv_in to
conv_out.  In that layout every 1x1 conv / Linear / LayerNorm is a row-wise op on ``[B*T*H*W, C]``,
the spatial transformer's NCHW<->NLC permutes (pnp_utils.py:434, :502) are views, and the temporal
transformer needs no ``[(b t) c h w] <-> [(b h w) t c]`` permute (pnp_utils.py:189, :207-213) at all:
its row-wise ops do not care about row order and the temporal attention kernel walks the frames of a
pixel with a stride.  The first NCHW version of this file spent ~40 % of a step in cuDNN
nchw<->nhwc conversions and strided elementwise copies (profiles/r01_launches_v1_nchw_summary.txt).

Forward bodies restate the functions MVOC re-points (i2vgen-xl/pnp_utils.py:170-548) and the stock
diffusers blocks; citations are on each method.
"""
from __future__ import annotations

import inspect
import math
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops



class UNetConfig:
    def __init__(self, block_out_channels=(320, 640, 1280, 1280),
                 down_block_types=("CrossAttnDownBlock3D",) * 3 + ("DownBlock3D",),
                 up_block_types=("UpBlock3D",) + ("CrossAttnUpBlock3D",) * 3,
                 layers_per_block=2, norm_num_groups=32, cross_attention_dim=1024, attention_head_dim=64,
                 in_channels=4, out_channels=4, transformer_in_heads=8):
        self.block_out_channels = tuple(block_out_channels)
        self.down_block_types = tuple(down_block_types)
        self.up_block_types = tuple(up_block_types)
        self.layers_per_block = layers_per_block
        self.norm_num_groups = norm_num_groups
        self.cross_attention_dim = cross_attention_dim
        self.attention_head_dim = attention_head_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.transformer_in_heads = transformer_in_heads

    @staticmethod
    def full():
        return UNetConfig()

    @staticmethod
    def reduced():
        return UNetConfig(block_out_channels=(64, 128), down_block_types=("CrossAttnDownBlock3D", "DownBlock3D"),
                          up_block_types=("UpBlock3D", "CrossAttnUpBlock3D"), transformer_in_heads=2)

    @staticmethod
    def tiny4():
        """Full 4-level layout with narrow channels (the golden-vector model of tests/golden)."""
        return UNetConfig(block_out_channels=(64, 128, 256, 256), transformer_in_heads=2)

    @staticmethod
    def named(kind: str):
        try:
            return {"full": UNetConfig.full, "reduced": UNetConfig.reduced, "tiny4": UNetConfig.tiny4}[kind]()
        except KeyError:
            raise ValueError(f"unknown UNet config {kind!r}") from None


# ------------------------------------------------------------------ small pieces
class _Ctx:
    """Per-UNet execution context shared by its modules: the (optional) multi-GPU partition."""

    def __init__(self):
        self.parallel = None   # mvoc_b200.parallel.FrameParallel or None
        self.full_hw = None    # (h, w) of the un-sharded frame while a temporal operator runs on pixel shards


# MVOC_STAGED=1 routes the stride-1 3x3 convolutions and the GEGLU projection through the tcgen05 kernels staged
# in libmvoc_b200_staged.so (include/mvoc_b200_staged.h).  Off by default: those kernels have not been validated
# on hardware yet, the measured product path is cuDNN / cuBLAS + mvoc_geglu.
_STAGED = os.environ.get("MVOC_STAGED") == "1"


def _staged_conv_ok(conv: nn.Conv2d) -> bool:
    return (conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.padding == (1, 1)
            and conv.dilation == (1, 1) and conv.groups == 1 and conv.in_channels % 64 == 0
            and conv.out_channels % 64 == 0)


def _staged_conv(x, conv, bias=None, residual=None):
    from . import staged

    wt = conv.__dict__.get("_w_taps")
    if wt is None or wt.device != x.device or wt.dtype != x.dtype:
        wt = staged.prepare_conv_weight(conv.weight.detach()).to(x.dtype)
        conv.__dict__["_w_taps"] = wt
    return staged.conv3x3_nhwc(x, wt, conv.bias if bias is None else bias, residual)


def conv_nhwc(x: torch.Tensor, conv: nn.Conv2d, bias: Optional[torch.Tensor] = None,
              residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 / strided conv on a channels-last activation [N, H, W, C] -> [N, H', W', C'] (cuDNN NHWC kernels:
    the permuted view IS torch's channels_last memory format, so no nchw<->nhwc conversion runs).
    `residual` (same shape as the result) is added to it."""
    if _STAGED and _staged_conv_ok(conv):
        return _staged_conv(x, conv, bias, residual)
    if residual is not None:
        return conv_nhwc(x, conv, bias).add_(residual)
    y = F.conv2d(x.permute(0, 3, 1, 2), conv.weight, conv.bias if bias is None else bias, conv.stride, conv.padding)
    y = y.permute(0, 2, 3, 1)
    return y if y.is_contiguous() else y.contiguous()


class GroupNormAct(nn.GroupNorm):
    """nn.GroupNorm parameters, executed by the channels-last GroupNorm kernels (optionally fused SiLU,
    fused per-(frame, channel) add, statistics over the T frames of a video)."""

    def forward(self, x, silu: bool = False, frames_per_stat: int = 1, add=None, out=None, gather=None):
        return ops.groupnorm_nhwc(x, self.weight, self.bias, self.num_groups, self.eps, silu, frames_per_stat,
                                  add, out, gather)


class Timesteps(nn.Module):
    def __init__(self, num_channels: int):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# ------------------------------------------------------------------ attention
class TemporalShape:
    """Marks hidden_states [B*T, S, C] as frame-major tokens whose attention runs over the T frames of
    each (video, pixel) — what the reference expresses by permuting to [(b h w), T, C] (pnp_utils.py:189)."""

    __slots__ = ("videos", "frames", "pixels")

    def __init__(self, videos: int, frames: int, pixels: int):
        self.videos, self.frames, self.pixels = videos, frames, pixels


def run_attention(q, k, v, heads, temporal: Optional[TemporalShape]):
    """q [N, Sq, C], k/v [N, Sk, C] (any row stride).  Spatial / cross: tcgen05 kernel over the S axis.
    Temporal: warp-per-(pixel, head) kernel over the frame axis, reading the frame-major rows in place."""
    if temporal is None:
        return ops.attention(q, k, v, heads)
    C = q.shape[-1]
    rows = q.shape[0] * q.shape[1]
    o = ops.temporal_attention_frames(q.view(rows, C), k.view(rows, C), v.view(rows, C), heads,
                                      temporal.videos, temporal.frames, temporal.pixels)
    return o.view(q.shape[0], q.shape[1], C)


class AttnProcessor2_0:
    """Stock processor (diffusers AttnProcessor2_0 == pnp_utils.py:576-612 + :674-704 minus the
    injection block) over the C-ABI attention kernels."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 scale: float = 1.0):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is never set on this path (pnp_utils.py:594)")
        if encoder_hidden_states is None:
            q, k, v = attn.qkv_self(hidden_states)
        else:
            q = attn.to_q(hidden_states)
            k, v = attn.kv_cross(encoder_hidden_states)
        out = run_attention(q, k, v, attn.heads, attn.temporal if encoder_hidden_states is None else None)
        return attn.to_out[0](out)


class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8,
                 dim_head: int = 64, bias: bool = False, out_bias: bool = True):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.heads = heads
        self.dim_head = dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.is_cross = cross_attention_dim is not None
        self.scale = dim_head ** -0.5
        self.rescale_output_factor = 1.0
        self.residual_connection = False
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(kv_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(kv_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self._processor = None
        self._accepts_hw = False
        self._w_qkv = None
        self._w_kv = None
        self.temporal: Optional[TemporalShape] = None   # set per call by the temporal transformer
        self.ctx: Optional[_Ctx] = None
        self.processor = AttnProcessor2_0()

    # `module.processor = obj` is how MVOC installs its processors (pnp_utils.py:715, :897)
    @property
    def processor(self):
        return self._processor

    @processor.setter
    def processor(self, p):
        object.__setattr__(self, "_processor", p)
        # attention_forward (pnp_utils.py:357-365) inspects the processor on every call; the
        # answer only changes when the processor does, so it is computed here once.
        object.__setattr__(self, "_accepts_hw", "height" in inspect.signature(p.__call__).parameters)

    def qkv_self(self, x):
        """One GEMM for the three self-attention projections; q, k, v are strided views."""
        if self._w_qkv is None or self._w_qkv.device != x.device or self._w_qkv.dtype != x.dtype:
            self._w_qkv = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0).contiguous()
        qkv = F.linear(x, self._w_qkv)
        c = self.inner_dim
        return qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]

    def kv_cross(self, ctx):
        if self._w_kv is None or self._w_kv.device != ctx.device or self._w_kv.dtype != ctx.dtype:
            self._w_kv = torch.cat([self.to_k.weight, self.to_v.weight], dim=0).contiguous()
        kv = F.linear(ctx, self._w_kv)
        c = self.inner_dim
        return kv[..., :c], kv[..., c:]

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, height=None, width=None,
                **cross_attention_kwargs):
        """attention_forward, pnp_utils.py:348-385."""
        if self._accepts_hw:
            return self._processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                                   attention_mask=attention_mask, height=height, width=width)
        return self._processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                               attention_mask=attention_mask)


class LayerNormK(nn.LayerNorm):
    """nn.LayerNorm parameters, executed by mvoc_layernorm (warp per row)."""

    def forward(self, x):
        return ops.layernorm(x, self.weight, self.bias, self.eps)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        if _STAGED and self.proj.in_features % 64 == 0 and self.proj.out_features % 128 == 0:
            from . import staged

            return staged.linear_geglu(x, self.proj.weight, self.proj.bias)
        return ops.geglu(self.proj(x))      # x * gelu(gate) in one pass (mvoc_geglu)


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4, activation_fn: str = "geglu"):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner) if activation_fn == "geglu" else GELU(dim, inner),
                                  nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    """basic_transformer_block_forward, pnp_utils.py:222-346 (norm_type 'layer_norm').  Row-wise except for
    the two attention cores; `temporal` selects the frame axis for them."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None,
                 double_self_attention=False):
        super().__init__()
        self.only_cross_attention = False
        self.norm_type = "layer_norm"
        self.norm1 = LayerNormK(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, num_attention_heads, attention_head_dim, bias=False)
        self.norm2 = LayerNormK(dim, eps=1e-5)
        self.attn2 = Attention(dim, None if double_self_attention else cross_attention_dim,
                               num_attention_heads, attention_head_dim, bias=False)
        self.norm3 = LayerNormK(dim, eps=1e-5)
        self.ff = FeedForward(dim, activation_fn="geglu")

    def forward(self, hidden_states, encoder_hidden_states=None, height=None, width=None,
                temporal: Optional[TemporalShape] = None):
        self.attn1.temporal = temporal
        self.attn2.temporal = temporal
        h = self.attn1(self.norm1(hidden_states), encoder_hidden_states=None, height=height, width=width)
        hidden_states = h.add_(hidden_states)                                   # :283
        h = self.attn2(self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states)
        hidden_states = h.add_(hidden_states)                                   # :315
        h = self.ff(self.norm3(hidden_states))                                  # :322-335
        return h.add_(hidden_states)                                            # :342


class Transformer2DModel(nn.Module):
    """transformer2dmodel_forward live branch, pnp_utils.py:426-434 and :462-508, on channels-last
    activations: the NCHW->NLC permute (:434) and its inverse (:502) are views and the two 1x1 convs are
    row-wise GEMMs."""

    def __init__(self, num_attention_heads, attention_head_dim, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.is_input_continuous, self.use_linear_projection = True, False
        self.norm = GroupNormAct(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, kernel_size=1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, kernel_size=1)

    def forward(self, hidden_states, encoder_hidden_states=None, **kwargs):
        n, height, width, C = hidden_states.shape
        inner = self.proj_in.out_channels
        x = self.norm(hidden_states)                                            # :430
        tokens = F.linear(x.view(n, height * width, C), self.proj_in.weight.view(inner, C), self.proj_in.bias)
        for block in self.transformer_blocks:                                   # :487-497
            tokens = block(tokens, encoder_hidden_states=encoder_hidden_states, height=height, width=width)
        out = F.linear(tokens, self.proj_out.weight.view(C, inner), self.proj_out.bias)   # :503
        out = out.view(n, height, width, C).add_(hidden_states)                 # :508
        return (out,)


class TransformerTemporalModel(nn.Module):
    """transformer_temporal_model_forward, pnp_utils.py:170-220, on frame-major channels-last rows."""

    def __init__(self, num_attention_heads, attention_head_dim, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = GroupNormAct(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim,
                                   double_self_attention=True)])
        self.proj_out = nn.Linear(inner, in_channels)
        self.ctx: Optional[_Ctx] = None

    def forward(self, hidden_states, encoder_hidden_states=None, num_frames: int = 1, **kwargs):
        par = self.ctx.parallel if self.ctx is not None else None
        if par is not None and par.world > 1:
            return (par.temporal_transformer(self, hidden_states, num_frames),)
        return (self.forward_local(hidden_states, num_frames, None),)

    def forward_local(self, hidden_states, num_frames, gather):
        """hidden_states [(b t), h, w, C] holding ALL frames of its pixels (h*w may be a pixel shard)."""
        bt, height, width, C = hidden_states.shape
        b, S = bt // num_frames, height * width
        x = self.norm(hidden_states, frames_per_stat=num_frames, gather=gather)  # 5-D GroupNorm, :185-188
        x = self.proj_in(x.view(bt, S, C))                                      # :191 (the permute at :189 is implicit)
        shape = TemporalShape(b, num_frames, S)
        for block in self.transformer_blocks:                                   # :194-203
            x = block(x, encoder_hidden_states=None, height=height, width=width, temporal=shape)
        x = self.proj_out(x)                                                    # :206
        return x.view(bt, height, width, C).add_(hidden_states)                 # :207-215


# ------------------------------------------------------------------ conv blocks
class ResnetBlock2D(nn.Module):
    """Stock ResnetBlock2D forward == the closure at pnp_utils.py:902-1020 without the injection block."""

    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = GroupNormAct(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = GroupNormAct(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.upsample = self.downsample = None
        self.output_scale_factor = 1.0
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None
        self.feature_hook = None  # set by register_resnet_injection
        self.ctx: Optional[_Ctx] = None
        self._bias2 = None

    def forward(self, input_tensor, temb, scale: float = 1.0):
        n, H, W, cin = input_tensor.shape
        h = self.norm1(input_tensor, silu=True)                                 # :909-910
        h = conv_nhwc(h, self.conv1)                                            # :939
        t = self.time_emb_proj(F.silu(temb))                                    # :941-948
        h = self.norm2(h, silu=True, add=t, out=h)                              # `+ temb` :952 fused into norm2 :953, :965
        if self.conv_shortcut is None:
            if self.feature_hook is None:
                return conv_nhwc(h, self.conv2, residual=input_tensor)          # :968 + :1018 (scale factor 1)
            h = conv_nhwc(h, self.conv2)                                        # :968
            self.feature_hook(self, h)                                          # :970-1004
            return h.add_(input_tensor)                                         # :1018
        # 1x1 shortcut conv (:1011-1016) == row GEMM accumulated into conv2's output; its bias rides on
        # conv2's (a per-channel constant commutes with the per-pixel select of the injection)
        if self._bias2 is None or self._bias2.device != h.device or self._bias2.dtype != h.dtype:
            self._bias2 = (self.conv2.bias + self.conv_shortcut.bias).detach()
        h = conv_nhwc(h, self.conv2, bias=self._bias2)
        if self.feature_hook is not None:
            self.feature_hook(self, h)
        cout = h.shape[-1]
        h.view(-1, cout).addmm_(input_tensor.view(-1, cin), self.conv_shortcut.weight.view(cout, cin).t())
        return h


class _TemporalTap(nn.Sequential):
    """[GroupNorm, SiLU, (Dropout), Conv3d(k=(3,1,1))] — an nn.Sequential so the state-dict names equal
    diffusers' ("conv1.0.weight", "conv1.2.weight", "conv2.3.weight", ...); executed as GroupNorm(5-D
    statistics)+SiLU in one kernel and the 3-tap temporal conv as accumulating row GEMMs on frame-shifted
    row ranges (no [B,C,T,H,W] permute and no im2col copy)."""

    def __init__(self, dim_in, dim_out, groups, with_dropout):
        mods = [GroupNormAct(groups, dim_in), nn.SiLU()]
        if with_dropout:
            mods.append(nn.Dropout(0.0))
        mods.append(nn.Conv3d(dim_in, dim_out, (3, 1, 1), padding=(1, 0, 0)))
        super().__init__(*mods)
        self._taps = None

    @property
    def norm(self):
        return self[0]

    @property
    def conv(self):
        return self[len(self) - 1]

    def forward(self, x, num_frames, gather=None):
        """x: [(b t), h, w, C] frame-major channels-last -> same layout."""
        bt, h, w, C = x.shape
        b, S = bt // num_frames, h * w
        y = self.norm(x, silu=True, frames_per_stat=num_frames, gather=gather)
        conv = self.conv
        co = conv.out_channels
        if self._taps is None or self._taps.device != x.device or self._taps.dtype != x.dtype:
            # [co, ci, 3, 1, 1] -> [tap, ci, co] (right-hand operands of the row GEMMs)
            self._taps = conv.weight.view(co, C, 3).permute(2, 1, 0).contiguous()
        wp, wc, wn = self._taps[0], self._taps[1], self._taps[2]
        y2 = y.view(bt * S, C)
        out = torch.addmm(conv.bias, y2, wc)                                    # centre tap, all frames
        rows = num_frames * S
        if num_frames > 1:
            for i in range(b):
                r0, r1 = i * rows, (i + 1) * rows
                out[r0 + S:r1].addmm_(y2[r0:r1 - S], wp)                        # tap on frame t-1
                out[r0:r1 - S].addmm_(y2[r0 + S:r1], wn)                        # tap on frame t+1
        return out.view(bt, h, w, co)


class TemporalConvLayer(nn.Module):
    """Stock TemporalConvLayer forward == the closure at pnp_utils.py:1042-1057 without the injection."""

    def __init__(self, in_dim, out_dim=None, dropout=0.1, norm_num_groups=32):
        super().__init__()
        out_dim = out_dim or in_dim
        self.conv1 = _TemporalTap(in_dim, out_dim, norm_num_groups, False)
        self.conv2 = _TemporalTap(out_dim, in_dim, norm_num_groups, True)
        self.conv3 = _TemporalTap(out_dim, in_dim, norm_num_groups, True)
        self.conv4 = _TemporalTap(out_dim, in_dim, norm_num_groups, True)
        nn.init.zeros_(self.conv4.conv.weight)
        nn.init.zeros_(self.conv4.conv.bias)
        self.feature_hook = None  # set by register_temp_conv_injection
        self.ctx: Optional[_Ctx] = None

    def forward(self, hidden_states, num_frames: int = 1):
        par = self.ctx.parallel if self.ctx is not None else None
        if par is not None and par.world > 1:
            h = par.temporal_conv(self, hidden_states, num_frames)
        else:
            h = self.forward_local(hidden_states, num_frames, None)
        if self.feature_hook is not None:
            self.feature_hook(self, h)                                          # :1059-1082
        return h

    def forward_local(self, hidden_states, num_frames, gather):
        h = self.conv1(hidden_states, num_frames, gather)                       # :1048
        h = self.conv2(h, num_frames, gather)
        h = self.conv3(h, num_frames, gather)
        h = self.conv4(h, num_frames, gather)                                   # :1051
        return h.add_(hidden_states)                                            # :1053


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x, scale: float = 1.0):
        return conv_nhwc(x, self.conv)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None, scale: float = 1.0):
        xc = x.permute(0, 3, 1, 2)                                              # channels_last view
        if output_size is None:
            xc = F.interpolate(xc, scale_factor=2.0, mode="nearest")
        else:
            xc = F.interpolate(xc, size=output_size, mode="nearest")
        y = F.conv2d(xc, self.conv.weight, self.conv.bias, 1, 1).permute(0, 2, 3, 1)
        return y if y.is_contiguous() else y.contiguous()


class _Block3D(nn.Module):
    has_cross_attention = False

    def _make(self, n, in_chs, out_ch, temb, head_dim, cross_dim, groups, attn):
        self.resnets = nn.ModuleList([ResnetBlock2D(ic, out_ch, temb, groups) for ic in in_chs])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(out_ch, out_ch, 0.1, groups) for _ in range(n)])
        if attn:
            heads = out_ch // head_dim
            self.attentions = nn.ModuleList(
                [Transformer2DModel(heads, head_dim, out_ch, cross_dim, groups) for _ in range(n)])
            self.temp_attentions = nn.ModuleList(
                [TransformerTemporalModel(heads, head_dim, out_ch, cross_dim, groups) for _ in range(n)])


class CrossAttnDownBlock3D(_Block3D):
    has_cross_attention = True

    def __init__(self, in_ch, out_ch, temb, n_layers, head_dim, cross_dim, groups, add_downsample):
        super().__init__()
        self._make(n_layers, [in_ch] + [out_ch] * (n_layers - 1), out_ch, temb, head_dim, cross_dim, groups, True)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, num_frames=1, **kw):
        outs = ()
        for resnet, temp_conv, attn, temp_attn in zip(self.resnets, self.temp_convs, self.attentions,
                                                      self.temp_attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states)[0]
            hidden_states = temp_attn(hidden_states, num_frames=num_frames)[0]
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class DownBlock3D(_Block3D):
    def __init__(self, in_ch, out_ch, temb, n_layers, groups, add_downsample):
        super().__init__()
        self._make(n_layers, [in_ch] + [out_ch] * (n_layers - 1), out_ch, temb, 0, 0, groups, False)
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, num_frames=1, **kw):
        outs = ()
        for resnet, temp_conv in zip(self.resnets, self.temp_convs):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class UNetMidBlock3DCrossAttn(_Block3D):
    has_cross_attention = True

    def __init__(self, ch, temb, head_dim, cross_dim, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups) for _ in range(2)])
        self.temp_convs = nn.ModuleList([TemporalConvLayer(ch, ch, 0.1, groups) for _ in range(2)])
        heads = ch // head_dim
        self.attentions = nn.ModuleList([Transformer2DModel(heads, head_dim, ch, cross_dim, groups)])
        self.temp_attentions = nn.ModuleList([TransformerTemporalModel(heads, head_dim, ch, cross_dim, groups)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, num_frames=1, **kw):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.temp_convs[0](hidden_states, num_frames=num_frames)
        for attn, temp_attn, resnet, temp_conv in zip(self.attentions, self.temp_attentions, self.resnets[1:],
                                                      self.temp_convs[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states)[0]
            hidden_states = temp_attn(hidden_states, num_frames=num_frames)[0]
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
        return hidden_states


def _up_in_channels(in_ch, out_ch, prev_out, n_layers):
    ins = []
    for i in range(n_layers):
        res_skip = in_ch if i == n_layers - 1 else out_ch
        res_in = prev_out if i == 0 else out_ch
        ins.append(res_in + res_skip)
    return ins


class CrossAttnUpBlock3D(_Block3D):
    has_cross_attention = True

    def __init__(self, in_ch, out_ch, prev_out, temb, n_layers, head_dim, cross_dim, groups, add_upsample):
        super().__init__()
        self._make(n_layers, _up_in_channels(in_ch, out_ch, prev_out, n_layers), out_ch, temb, head_dim, cross_dim,
                   groups, True)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                upsample_size=None, num_frames=1, **kw):
        for resnet, temp_conv, attn, temp_attn in zip(self.resnets, self.temp_convs, self.attentions,
                                                      self.temp_attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=-1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states)[0]
            hidden_states = temp_attn(hidden_states, num_frames=num_frames)[0]
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class UpBlock3D(_Block3D):
    def __init__(self, in_ch, out_ch, prev_out, temb, n_layers, groups, add_upsample):
        super().__init__()
        self._make(n_layers, _up_in_channels(in_ch, out_ch, prev_out, n_layers), out_ch, temb, 0, 0, groups, False)
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, num_frames=1, **kw):
        for resnet, temp_conv in zip(self.resnets, self.temp_convs):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=-1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = temp_conv(hidden_states, num_frames=num_frames)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class TinyAttention(nn.Module):
    """Attention(query_dim=4, heads=2, dim_head=4) of I2VGenXLTransformerTemporalEncoder: head_dim 4 is
    below any tensor-core tile; 0.0002 % of the FLOPs.  Computed with explicit matmuls (no SDPA library)."""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim, bias=True), nn.Dropout(0.0)])

    def forward(self, x):
        B, N, _ = x.shape
        q = self.to_q(x).view(B, N, self.heads, self.dim_head).transpose(1, 2).float()
        k = self.to_k(x).view(B, N, self.heads, self.dim_head).transpose(1, 2).float()
        v = self.to_v(x).view(B, N, self.heads, self.dim_head).transpose(1, 2).float()
        p = torch.softmax(q @ k.transpose(-1, -2) * (self.dim_head ** -0.5), dim=-1)
        o = (p @ v).transpose(1, 2).reshape(B, N, self.heads * self.dim_head).to(x.dtype)
        return self.to_out[0](o)


class I2VGenXLTransformerTemporalEncoder(nn.Module):
    def __init__(self, dim, num_attention_heads=2, attention_head_dim=4):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = TinyAttention(dim, num_attention_heads, attention_head_dim)
        self.ff = FeedForward(dim, activation_fn="gelu")

    def forward(self, hidden_states):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        return self.ff(hidden_states) + hidden_states


class _Cfg:
    def __init__(self, c):
        self.in_channels = c.in_channels
        self.cross_attention_dim = c.cross_attention_dim


class I2VGenXLUNet(nn.Module):
    def __init__(self, cfg: Optional[UNetConfig] = None):
        super().__init__()
        cfg = cfg or UNetConfig()
        self.cfg = cfg
        self.config = _Cfg(cfg)
        boc = cfg.block_out_channels
        temb = boc[0] * 4
        g, hd, cd, ic = cfg.norm_num_groups, cfg.attention_head_dim, cfg.cross_attention_dim, cfg.in_channels
        self.conv_in = nn.Conv2d(ic + ic, boc[0], 3, padding=1)
        self.transformer_in = TransformerTemporalModel(cfg.transformer_in_heads, hd, boc[0], cd, g)
        self.image_latents_proj_in = nn.Sequential(
            nn.Conv2d(4, ic * 4, 3, padding=1), nn.SiLU(), nn.Conv2d(ic * 4, ic * 4, 3, padding=1), nn.SiLU(),
            nn.Conv2d(ic * 4, ic, 3, padding=1))
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
            if not final:
                self.num_upsamplers += 1
            if t == "CrossAttnUpBlock3D":
                self.up_blocks.append(CrossAttnUpBlock3D(in_ch, out_ch, prev_out, temb, cfg.layers_per_block + 1,
                                                         hd, cd, g, not final))
            else:
                self.up_blocks.append(UpBlock3D(in_ch, out_ch, prev_out, temb, cfg.layers_per_block + 1, g,
                                                not final))
        self.conv_norm_out = GroupNormAct(g, boc[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)
        self.conv_out.feature_hook = None  # set by register_out_conv_injection
        self.ctx = _Ctx()
        for m in self.modules():
            if isinstance(m, (TransformerTemporalModel, TemporalConvLayer, Attention, ResnetBlock2D)):
                m.ctx = self.ctx

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    # ------------------------------------------------------------------
    def _embeddings(self, sample, timestep, fps):
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.int64, device=sample.device)
        elif timesteps.dim() == 0:
            timesteps = timesteps[None].to(sample.device)
        timesteps = timesteps.expand(sample.shape[0])
        t_emb = self.time_embedding(self.time_proj(timesteps).to(self.dtype))
        fps_emb = self.fps_embedding(self.time_proj(fps.expand(fps.shape[0])).to(self.dtype))
        return t_emb + fps_emb

    def context(self, encoder_hidden_states, image_latents, image_embeddings):
        """Context tokens [B, 77 + 64 + 4, 1024] of ONE frame per video (pipeline_i2vgen_xl.py:204-240 with
        multi_frame_guidance False computes the same tensor T times; it only depends on frame 0)."""
        il = image_latents[:, :, 0]
        il = self.image_latents_context_embedding(il)
        b, c, h, w = il.shape
        il = il.permute(0, 2, 3, 1).reshape(b, h * w, c)
        image_emb = self.context_embedding(image_embeddings[:, 0:1, :])
        image_emb = image_emb.view(-1, self.config.in_channels, self.config.cross_attention_dim)
        return torch.cat([encoder_hidden_states, il, image_emb], dim=1)

    def stem_condition(self, image_latents_first):
        """Image-latent conditioning channels [(b t), c, h, w] (pipeline_i2vgen_xl.py:264-279); independent
        of the timestep, so the step loop computes it once."""
        b, c, T, h, w = image_latents_first.shape
        il = image_latents_first.permute(0, 2, 1, 3, 4).reshape(b * T, c, h, w)
        il = self.image_latents_proj_in(il)
        il = il.view(b, T, c, h, w).permute(0, 3, 4, 1, 2).reshape(b * h * w, T, c)
        il = self.image_latents_temporal_encoder(il)
        return il.reshape(b, h, w, T, c).permute(0, 3, 4, 1, 2).reshape(b * T, c, h, w).contiguous()

    def stem(self, sample_frames, il_frames, num_frames):
        """cat + conv_in + transformer_in (pipeline_i2vgen_xl.py:282-290).  sample_frames, il_frames:
        [(b t), c, h, w] (the frames this rank owns); returns channels-last [(b t), h, w, C0]."""
        x = torch.cat([sample_frames, il_frames], dim=1).permute(0, 2, 3, 1).contiguous()
        x = conv_nhwc(x, self.conv_in)
        return self.transformer_in(x, num_frames=num_frames)[0]

    def body(self, sample, emb, context_emb, num_frames, forward_upsample_size):
        """down / mid / up / out, pipeline_i2vgen_xl.py:292-357, on channels-last [(b t), h, w, C].
        Returns the noise prediction as [(b t), 4, h, w] (contiguous NCHW)."""
        upsample_size = None
        res_samples = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, res = blk(sample, temb=emb, encoder_hidden_states=context_emb, num_frames=num_frames)
            else:
                sample, res = blk(sample, temb=emb, num_frames=num_frames)
            res_samples += res
        sample = self.mid_block(sample, emb, encoder_hidden_states=context_emb, num_frames=num_frames)
        for i, blk in enumerate(self.up_blocks):
            n = len(blk.resnets)
            res = res_samples[-n:]
            res_samples = res_samples[:-n]
            if i != len(self.up_blocks) - 1 and forward_upsample_size:
                upsample_size = res_samples[-1].shape[1:3]
            if blk.has_cross_attention:
                sample = blk(sample, res, temb=emb, encoder_hidden_states=context_emb, upsample_size=upsample_size,
                             num_frames=num_frames)
            else:
                sample = blk(sample, res, temb=emb, upsample_size=upsample_size, num_frames=num_frames)
        sample = self.conv_norm_out(sample, silu=True, out=sample)              # :351-352
        sample = conv_nhwc(sample, self.conv_out).permute(0, 3, 1, 2).contiguous()   # :354 -> [(b t), 4, h, w]
        if self.conv_out.feature_hook is not None:
            self.conv_out.feature_hook(self.conv_out, sample)                   # pnp_utils.py:1114-1146
        return sample

    def forward(self, sample, timestep, fps, image_latents, image_embeddings=None, encoder_hidden_states=None,
                image_latents_first=None, return_dict: bool = False, **kwargs):
        """Stock signature (diffusers I2VGenXLUNet.forward) plus MVOC's `image_latents_first`
        (I2VGenXLUnetExtension.forward, pipeline_i2vgen_xl.py:109-122).  Single-GPU convenience entry; the
        step loops use the pieces directly (mvoc_b200/pipeline.py).  Returns a 1-tuple [b, 4, T, h, w]."""
        if not sample.is_cuda:
            raise RuntimeError("mvoc_b200.I2VGenXLUNet runs on CUDA tensors only (no CPU fallback)")
        if image_latents_first is None:
            image_latents_first = image_latents
        b, c, T, h, w = sample.shape
        fwd_up = any(s % (2 ** self.num_upsamplers) != 0 for s in (h, w))
        emb = self._embeddings(sample, timestep, fps).repeat_interleave(T, dim=0)
        ctx = self.context(encoder_hidden_states, image_latents, image_embeddings)
        ctx = ctx.repeat_interleave(T, dim=0)
        frames = sample.permute(0, 2, 1, 3, 4).reshape(b * T, c, h, w)
        x = self.stem(frames, self.stem_condition(image_latents_first), T)
        out = self.body(x, emb, ctx, T, fwd_up)
        return (out.view(b, T, *out.shape[1:]).permute(0, 2, 1, 3, 4),)


def build_unet(kind: str = "full", seed: int = 0, device="cuda", dtype=torch.bfloat16) -> I2VGenXLUNet:
    """Random-init weights of the named architecture: torch.manual_seed(seed) + PyTorch default inits on the
    CPU in fp32, then cast/moved.  (Parity tests do not rely on RNG order: they load the oracle's state dict.)"""
    torch.manual_seed(seed)
    m = I2VGenXLUNet(UNetConfig.named(kind)).eval().requires_grad_(False)
    return prepare(m.to(device=device, dtype=dtype))


def prepare(unet: I2VGenXLUNet) -> I2VGenXLUNet:
    """Put the conv weights in channels_last memory format once (cuDNN then runs its NHWC kernels with no
    per-call filter transform)."""
    for m in unet.modules():
        if isinstance(m, nn.Conv2d) and m.kernel_size != (1, 1):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    return unet
