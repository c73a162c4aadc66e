"""Bindings of the kernels STAGED for the next round (include/mvoc_b200_staged.h, libmvoc_b200_staged.so).

They compile for sm_100a but have not run on hardware yet; nothing in the product imports this module unless
MVOC_STAGED=1 is set.  Same rules as mvoc_b200.ops: CUDA tensors only, no fallback, errors raise."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_void_p
from typing import Optional

import torch

from . import _cabi

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libmvoc_b200_staged.so")

SIGNATURES = {
    "mvoc_conv3x3_nhwc": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    "mvoc_linear_geglu": (c_int, [c_void_p] * 4 + [c_int64, c_int, c_int, c_int, c_void_p]),
    "mvoc_attn_fwd_split": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_int64] * 12 + [c_float, c_int, c_int, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise _cabi.MvocLibraryError(f"{LIB_PATH} not found: run __graft_entry__.build(); there is no fallback")
        lib = ctypes.CDLL(LIB_PATH)
        lib.mvoc_last_error.restype = ctypes.c_char_p
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {load().mvoc_last_error().decode()}")
    from . import ops

    ops._count()          # bench.py's gpu_launches counts these launches too


def _need(*ts) -> None:
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous()):
            raise ValueError("staged kernels take contiguous bf16 CUDA tensors (no fallback)")


def prepare_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """torch Conv2d weight [Cout, Cin, 3, 3] -> tap-major [9, Cout, Cin] (tap = kh*3 + kw), contiguous."""
    co, ci, kh, kw = weight.shape
    if (kh, kw) != (3, 3):
        raise ValueError("3x3 kernels only")
    return weight.permute(2, 3, 0, 1).reshape(9, co, ci).contiguous()


def conv3x3_nhwc(x: torch.Tensor, w_taps: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                 variant: int = 0) -> torch.Tensor:
    """x [N, H, W, Cin], w_taps [9, Cout, Cin] -> [N, H, W, Cout] (+ bias, + residual), stride 1, padding 1."""
    _need(x, w_taps, bias, residual, out)
    N, H, W, ci = x.shape
    if w_taps.dim() != 3 or w_taps.shape[0] != 9 or w_taps.shape[2] != ci:
        raise ValueError(f"w_taps must be [9, Cout, {ci}], got {tuple(w_taps.shape)}")
    co = w_taps.shape[1]
    if out is None:
        out = torch.empty((N, H, W, co), dtype=x.dtype, device=x.device)
    if residual is not None and residual.shape != out.shape:
        raise ValueError("residual must have the shape of the output")
    p = lambda t: None if t is None else t.data_ptr()
    rc = load().mvoc_conv3x3_nhwc(x.data_ptr(), w_taps.data_ptr(), p(bias), p(residual), out.data_ptr(), N, H, W, ci,
                                  co, _cabi.MVOC_BF16, int(variant), torch.cuda.current_stream().cuda_stream)
    _check(rc, "mvoc_conv3x3_nhwc")
    return out


def linear_geglu(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [..., K], weight [2F, K] (torch Linear) -> value * gelu(gate), [..., F]."""
    _need(x, weight, bias, out)
    K = x.shape[-1]
    M = x.numel() // K
    F2 = weight.shape[0]
    if weight.shape[1] != K or F2 % 2:
        raise ValueError(f"weight must be [2F, {K}], got {tuple(weight.shape)}")
    F = F2 // 2
    if out is None:
        out = torch.empty(x.shape[:-1] + (F,), dtype=x.dtype, device=x.device)
    rc = load().mvoc_linear_geglu(x.data_ptr(), weight.data_ptr(), None if bias is None else bias.data_ptr(),
                                  out.data_ptr(), M, K, F, _cabi.MVOC_BF16, torch.cuda.current_stream().cuda_stream)
    _check(rc, "mvoc_linear_geglu")
    return out


def attention_split(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
                    out: Optional[torch.Tensor] = None, variant: int = 0) -> torch.Tensor:
    """ops.attention through the row-split kernel: q [B,Nq,H*64], k/v [B,Nk,H*64] (last dim contiguous)."""
    for t in (q, k, v, out):
        if t is not None and (not t.is_cuda or t.dtype != torch.bfloat16 or t.stride(-1) != 1):
            raise ValueError("attention_split takes bf16 CUDA tensors with a contiguous last dim (no fallback)")
    B, Nq, C = q.shape
    if C != heads * 64 or k.shape[0] != B or v.shape != k.shape:
        raise ValueError(f"attention_split: q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)} heads {heads}")
    if out is None:
        out = torch.empty((B, Nq, C), dtype=q.dtype, device=q.device)
    st = lambda t: (t.stride(0), t.stride(1), 64)
    rc = load().mvoc_attn_fwd_split(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, heads, Nq, k.shape[1],
                                    64, *st(q), *st(k), *st(v), *st(out), float(scale if scale is not None else 0.125),
                                    _cabi.MVOC_BF16, int(variant), torch.cuda.current_stream().cuda_stream)
    _check(rc, "mvoc_attn_fwd_split")
    return out
