"""Kernels staged for the next round (include/mvoc_b200_staged.h).  CPU: the separate library loads, exports what
its header declares, and validates arguments.  The numerics tests run only on a GPU with MVOC_STAGED=1 — they are
NOT part of the `-m gpu` suite because the kernels have not run on hardware yet."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_staged_gpu = pytest.mark.skipif(os.environ.get("MVOC_STAGED") != "1" or not torch.cuda.is_available(),
                                      reason="staged kernels: set MVOC_STAGED=1 on a B200")


def test_staged_library_exports_its_header():
    from mvoc_b200 import staged

    hdr = open(os.path.join(ROOT, "include", "mvoc_b200_staged.h")).read()
    declared = set(re.findall(r"\b(mvoc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(staged.SIGNATURES)
    lib = staged.load()
    for name in declared:
        assert hasattr(lib, name)


def test_staged_argument_errors():
    from mvoc_b200 import staged

    lib = staged.load()
    assert lib.mvoc_attn_fwd_split(None, None, None, None, 1, 1, 128, 128, 64, *([0] * 12), 0.125, 0, 0, None) == -1


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@needs_staged_gpu
@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 5, 256, 256), (1, 1, 128, 128), (2, 2, 64, 64), (3, 10, 1024, 1024),
                                       (2, 5, 256, 145), (1, 3, 200, 77), (1, 5, 4096, 4096), (1, 2, 880, 880),
                                       (1, 2, 300, 200)])
def test_attention_split_vs_oracle_and_product(B, H, Nq, Nk, variant):
    from mvoc_b200 import ops, staged
    from oracle import ops_ref

    torch.manual_seed(B * 1000 + Nq + Nk)
    C = H * 64
    q, k, v = torch.randn(B, Nq, C).bfloat16(), torch.randn(B, Nk, C).bfloat16(), torch.randn(B, Nk, C).bfloat16()
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    out = staged.attention_split(q.cuda(), k.cuda(), v.cuda(), H, variant=variant)
    prod = ops.attention(q.cuda(), k.cuda(), v.cuda(), H)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert _rel(out.cpu(), ref) <= 1e-2
    assert _rel(out, prod) <= 5e-3


@needs_staged_gpu
def test_attention_split_large_scores():
    """Rescale path: scores grow along the key axis, so the running maximum keeps moving by more than 2^8."""
    from mvoc_b200 import staged
    from oracle import ops_ref

    torch.manual_seed(3)
    B, H, N = 1, 2, 1024
    q = torch.randn(B, N, H * 64).bfloat16()
    k = (torch.randn(B, N, H * 64) * torch.linspace(0.2, 6.0, N)[None, :, None]).bfloat16()
    v = torch.randn(B, N, H * 64).bfloat16()
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    out = staged.attention_split(q.cuda(), k.cuda(), v.cuda(), H)
    torch.cuda.synchronize()
    assert _rel(out.cpu(), ref) <= 2e-2
