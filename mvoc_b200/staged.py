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
