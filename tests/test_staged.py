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
    assert lib.mvoc_conv3x3_nhwc(None, None, None, None, None, 1, 8, 8, 64, 64, 0, 0, None) == -1
    assert b"null pointer" in lib.mvoc_last_error()
    assert lib.mvoc_conv3x3_nhwc(16, 16, None, None, 16, 1, 8, 8, 48, 64, 0, 0, None) == -2
    assert b"multiples of 64" in lib.mvoc_last_error()
    assert lib.mvoc_conv3x3_nhwc(16, 16, None, None, 16, 1, 8, 8, 64, 64, 1, 0, None) == -2      # fp16
    assert lib.mvoc_linear_geglu(16, 16, None, 16, 128, 100, 64, 0, None) == -2
    assert b"K=100" in lib.mvoc_last_error()
    w = torch.randn(6, 5, 3, 3)
    wt = staged.prepare_conv_weight(w)
    assert wt.shape == (9, 6, 5) and torch.equal(wt[1 * 3 + 2], w[:, :, 1, 2])


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@needs_staged_gpu
@pytest.mark.parametrize("shape", [
    (2, 64, 64, 64, 64),      # N, H, W, Cin, Cout: 2x64 box, BN 64
    (4, 32, 32, 128, 128),    # 4x32 box, BN 128
    (16, 8, 8, 128, 160),     # 2 frames per box, BN 160
    (3, 16, 16, 64, 320),     # BN 320 (two MMA pieces) and BN 160 via variant 1; N not a multiple of the box
    (5, 16, 16, 128, 640),    # two column tiles; 10 pixel tiles -> 5 CTA pairs per column tile (variant 2)
    (3, 8, 8, 64, 320),       # 2 frames per box: 2 pixel tiles = one CTA pair (variant 2)
    (5, 8, 8, 64, 320),       # 3 pixel tiles: the CTA-pair variant pads with one all-out-of-range tile
    (2, 11, 20, 64, 64),      # ragged H, W (config 5's lowest level): zero-filled pixels, masked stores
])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_conv3x3_vs_torch(shape, variant):
    from mvoc_b200 import staged

    N, H, W, ci, co = shape
    if variant == 2 and co % 320:
        pytest.skip("the CTA-pair variant needs Cout % 320 == 0")
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, H, W, ci, device="cuda", generator=g).bfloat16()
    w = (torch.randn(co, ci, 3, 3, device="cuda", generator=g) * (9 * ci) ** -0.5).bfloat16()
    b = torch.randn(co, device="cuda", generator=g).bfloat16()
    res = torch.randn(N, H, W, co, device="cuda", generator=g).bfloat16()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    out = staged.conv3x3_nhwc(x, staged.prepare_conv_weight(w), b, None, variant=variant)
    out_r = staged.conv3x3_nhwc(x, staged.prepare_conv_weight(w), b, res, variant=variant)
    torch.cuda.synchronize()
    assert _rel(out, ref) <= 5e-3
    assert _rel(out_r, ref + res.float()) <= 5e-3


@needs_staged_gpu
@pytest.mark.parametrize("M,K,F", [(256, 64, 64), (1000, 320, 1280), (4096, 640, 2560), (300, 128, 192)])
def test_linear_geglu_vs_torch(M, K, F):
    from mvoc_b200 import staged

    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(2 * F, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(2 * F, device="cuda", generator=g).bfloat16()
    y = torch.nn.functional.linear(x.float(), w.float(), b.float())
    ref = y[:, :F] * torch.nn.functional.gelu(y[:, F:])
    out = staged.linear_geglu(x, w, b)
    torch.cuda.synchronize()
    assert _rel(out, ref) <= 5e-3


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


def _choose_box(N, H, W):
    """Python mirror of gemm::choose_box (gemm_tc.cu): powers of two (bn, bh, bw), product 128, least padding."""
    best = None
    w = 128
    while w >= 1:
        h = 128 // w
        while h >= 1:
            n = 128 // (w * h)
            padded = -(-W // w) * w * (-(-H // h) * h) * (-(-N // n) * n)
            if best is None or padded < best[0]:
                best = (padded, n, h, w)
            h //= 2
        w //= 2
    return best[1:]


@pytest.mark.parametrize("N,H,W,ci,co", [(2, 11, 20, 64, 64), (3, 16, 16, 128, 64), (5, 8, 8, 64, 128)])
def test_conv_tiling_model(N, H, W, ci, co):
    """CPU model of the staged conv kernel's data movement (not of the hardware): per CTA, nine shifted, zero-filled
    boxes {64 ch, bw, bh, bn} of the NHWC activation times the tap-major weights, rows = (n, h, w) of the box,
    stores masked to real pixels — must equal conv2d.  Guards the tile / tap / layout arithmetic of gemm_tc.cu."""
    from mvoc_b200 import staged

    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, H, W, ci, generator=g)
    w = torch.randn(co, ci, 3, 3, generator=g)
    wt = staged.prepare_conv_weight(w)                                   # [9, co, ci]
    bn, bh, bw = _choose_box(N, H, W)
    assert bn * bh * bw == 128
    out = torch.full((N, H, W, co), float("nan"))

    def box(n0, h0, w0, c0):                                             # TMA tile load with zero OOB fill
        t = torch.zeros(bn, bh, bw, 64)
        for a in range(bn):
            for b in range(bh):
                for c in range(bw):
                    n, hh, ww = n0 + a, h0 + b, w0 + c
                    if 0 <= n < N and 0 <= hh < H and 0 <= ww < W:
                        t[a, b, c] = x[n, hh, ww, c0:c0 + 64]
        return t.reshape(128, 64)

    for tn in range(-(-N // bn)):
        for th in range(-(-H // bh)):
            for tw in range(-(-W // bw)):
                acc = torch.zeros(128, co)
                for kc in range(ci // 64):
                    for tap in range(9):
                        kh, kw = divmod(tap, 3)
                        a = box(tn * bn, th * bh + kh - 1, tw * bw + kw - 1, kc * 64)
                        acc += a @ wt[tap, :, kc * 64:(kc + 1) * 64].t()
                for row in range(128):
                    iw, ih, i_n = row % bw, (row // bw) % bh, row // (bw * bh)
                    n, hh, ww = tn * bn + i_n, th * bh + ih, tw * bw + iw
                    if n < N and hh < H and ww < W:
                        out[n, hh, ww] = acc[row]
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    assert not torch.isnan(out).any()
    assert torch.allclose(out, ref, atol=1e-3, rtol=1e-4)
