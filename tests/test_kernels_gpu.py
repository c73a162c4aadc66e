"""GPU parity tests: every C-ABI kernel against the fp32 oracle op on identical bf16-rounded inputs.

Bars: binary-mask blends bit-exact; GN / soft blends / DDIM relative L2 <= 3e-3
(one bf16 rounding); attention relative L2 <= 1e-2 (bf16 P in the P.V product).
"""
import math

import pytest
import torch

from oracle import ops_ref

pytestmark = pytest.mark.gpu


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.float().cpu()
    b = b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def make_masks(n_obj, T, H, W, seed=0, device="cpu"):
    """Moving soft-edged ellipses (float in [0,1], bool = float > 10/255) shaped [1,4,T,H,W]."""
    g = torch.Generator().manual_seed(seed)
    ys = torch.arange(H).view(1, H, 1).float()
    xs = torch.arange(W).view(1, 1, W).float()
    out = []
    for j in range(n_obj):
        cy = H * (0.3 + 0.4 * torch.rand(1, generator=g)) + torch.linspace(0, H * 0.1, T).view(T, 1, 1)
        cx = W * (0.2 + 0.6 * j / max(1, n_obj)) + torch.linspace(0, W * 0.15, T).view(T, 1, 1)
        ry, rx = H * 0.18, W * 0.14
        d = torch.sqrt(((ys - cy) / ry) ** 2 + ((xs - cx) / rx) ** 2)
        edge = 2.0 / min(ry, rx)
        mf = ((1.0 + edge - d) / edge).clamp(0, 1)
        mf = (mf * 255).round() / 255
        mb = mf > (10.0 / 255.0)
        mf5 = mf[None, None].expand(1, 4, T, H, W).contiguous().to(device)
        mb5 = mb[None, None].expand(1, 4, T, H, W).contiguous().to(device)
        out.append((mf5, mb5))
    return out


# ------------------------------------------------------------------ attention
ATTN_CASES = [
    # B, H, Nq, Nk
    (2, 5, 256, 256),
    (1, 1, 128, 128),
    (2, 2, 64, 64),      # 8x8 latents (l3)
    (3, 10, 1024, 1024),
    (2, 5, 256, 145),    # cross attention: ragged Nk
    (1, 3, 200, 77),     # ragged Nq and Nk
    (1, 5, 4096, 4096),  # l0 spatial
    (1, 2, 880, 880),    # 22x40 (config 5 l2): not a tile multiple
]


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("B,H,Nq,Nk", ATTN_CASES)
def test_attention_vs_oracle(cuda_device, B, H, Nq, Nk, variant):
    from mvoc_b200 import ops

    torch.manual_seed(B * 1000 + Nq + Nk)
    C = H * 64
    q = torch.randn(B, Nq, C).bfloat16()
    k = torch.randn(B, Nk, C).bfloat16()
    v = torch.randn(B, Nk, C).bfloat16()
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    out = ops.attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device), H, variant=variant)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    err = rel_l2(out, ref)
    assert err <= 1e-2, f"rel L2 {err:.3e} (variant {variant})"


@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 5, 256, 256), (2, 5, 256, 145), (1, 3, 200, 77), (1, 5, 4096, 4096)])
def test_attention_fp16_vs_oracle(cuda_device, B, H, Nq, Nk):
    """fp16 storage (the reference's dtype on GPU, composite.py:76-77): same kernel, fp16 operand formats."""
    from mvoc_b200 import ops

    torch.manual_seed(B * 1000 + Nq + Nk)
    C = H * 64
    q, k, v = torch.randn(B, Nq, C).half(), torch.randn(B, Nk, C).half(), torch.randn(B, Nk, C).half()
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    out = ops.attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device), H)
    torch.cuda.synchronize()
    assert out.dtype == torch.float16 and torch.isfinite(out.float()).all()
    err = rel_l2(out, ref)
    assert err <= 3e-3, f"rel L2 {err:.3e}"     # fp16 P carries 3 more mantissa bits than bf16 P


@pytest.mark.parametrize("variant", [2, 4])
@pytest.mark.parametrize("B,H,N,pair,dtype", [(4, 2, 256, 4, torch.bfloat16), (2, 5, 1024, 2, torch.bfloat16),
                                              (3, 1, 200, 3, torch.bfloat16), (1, 5, 4096, 1, torch.bfloat16),
                                              (2, 2, 320, 2, torch.float16)])
def test_attention_pair_vs_oracle(cuda_device, B, H, N, pair, dtype, variant):
    """One softmax, two P.V products (mvoc_attn_pair_fwd): both outputs against the fp32 oracle, and the first
    output against the plain kernel (same math, 64-key instead of 128-key blocks)."""
    from mvoc_b200 import ops

    torch.manual_seed(N + B)
    C = H * 64
    q, k = torch.randn(B, N, C).to(dtype), torch.randn(B, N, C).to(dtype)
    v = torch.randn(B + pair, N, C).to(dtype)
    qd, kd, vd = q.to(cuda_device), k.to(cuda_device), v.to(cuda_device)
    out = ops.attention_pair(qd, kd, vd, H, pair, variant=variant)
    torch.cuda.synchronize()
    ref0 = ops_ref.sdpa_ref(q.float(), k.float(), v[:B].float(), H)
    ref1 = ops_ref.sdpa_ref(q.float(), k.float(), v[pair:].float(), H)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out[:B], ref0) <= 1e-2 and rel_l2(out[pair:], ref1) <= 1e-2
    plain = ops.attention(qd, kd, vd[:B].contiguous(), H, variant=variant)
    assert rel_l2(out[:B], plain.float()) <= 5e-3


def test_attention_large_logits_and_strided(cuda_device):
    """Peaked softmax (exercises the lazy O rescale) on q/k/v that are slices of one fused buffer."""
    from mvoc_b200 import ops

    torch.manual_seed(7)
    B, H, N = 2, 5, 1024
    C = H * 64
    qkv = (torch.randn(B, N, 3 * C) * 3.0).bfloat16()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    # make later keys much larger so the running max keeps growing
    ramp = torch.linspace(0.2, 4.0, N).view(1, N, 1)
    k = (k.float() * ramp).bfloat16()
    qkv_d = torch.cat([q, k, v], dim=-1).to(cuda_device)
    qd, kd, vd = qkv_d[..., :C], qkv_d[..., C:2 * C], qkv_d[..., 2 * C:]
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    for variant in (1, 2, 3, 4, 5):
        out = ops.attention(qd, kd, vd, H, variant=variant)
        torch.cuda.synchronize()
        err = rel_l2(out, ref)
        assert err <= 1e-2, f"rel L2 {err:.3e} (variant {variant})"


def test_attention_rejects_bad_shapes(cuda_device):
    from mvoc_b200 import _cabi, ops

    q = torch.randn(1, 128, 96, device=cuda_device).bfloat16()
    with pytest.raises((ValueError, _cabi.MvocError)):
        ops.attention(q, q, q, heads=2)  # head_dim 48
    q32 = torch.randn(1, 128, 64, device=cuda_device)
    with pytest.raises(_cabi.MvocError):
        ops.attention(q32, q32, q32, heads=1)  # fp32 unsupported, no fallback


@pytest.mark.parametrize("P,T,H", [(1024, 16, 5), (300, 16, 10), (257, 8, 2), (64, 32, 20), (50, 24, 1), (33, 4, 2),
                                   (17, 2, 1), (9, 19, 3)])
def test_temporal_attention_vs_oracle(cuda_device, P, T, H):
    from mvoc_b200 import ops

    torch.manual_seed(P + T)
    C = H * 64
    q = torch.randn(P, T, C).bfloat16()
    k = torch.randn(P, T, C).bfloat16()
    v = torch.randn(P, T, C).bfloat16()
    ref = ops_ref.sdpa_ref(q.float(), k.float(), v.float(), H)
    out = ops.temporal_attention(q.to(cuda_device), k.to(cuda_device), v.to(cuda_device), H)
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    assert err <= 1e-2, f"rel L2 {err:.3e}"


# ------------------------------------------------------------------ blends
def _token_masks_spatial(masks, h, w):
    """[n_obj, T*h*w] uint8 in (frame, pixel) order from the bool masks (nearest-resized)."""
    rows = []
    for _, mb in masks:
        m = ops_ref.nearest_mask(mb[0, 0].float(), h, w)  # [T,h,w]
        rows.append((m > 0.5).to(torch.uint8).reshape(-1))
    return torch.stack(rows).contiguous()


def _token_masks_temporal(masks, h, w):
    """[n_obj, h*w*T] float32 in (pixel, frame) order from the float masks (nearest-resized)."""
    rows = []
    for mf, _ in masks:
        m = ops_ref.nearest_mask(mf[0, 0].float(), h, w)  # [T,h,w]
        rows.append(m.permute(1, 2, 0).reshape(-1).float())
    return torch.stack(rows).contiguous()


@pytest.mark.parametrize("n_obj", [1, 2, 3])
@pytest.mark.parametrize("inject_background", [False, True])
def test_spatial_qk_blend_bit_exact(cuda_device, n_obj, inject_background):
    from mvoc_b200 import ops

    T, H, W, h, w, C = 8, 32, 32, 16, 16, 128
    nb = n_obj + 3
    masks = make_masks(n_obj, T, H, W, seed=n_obj)
    torch.manual_seed(3)
    q = torch.randn(nb * T, h * w, C).bfloat16()
    k = torch.randn(nb * T, h * w, C).bfloat16()
    q_ref, k_ref = ops_ref.spatial_qk_inject_ref(q.float(), k.float(), masks, h, w, inject_background)
    qd, kd = q.to(cuda_device), k.to(cuda_device)
    ops.qk_blend_(qd, kd, _token_masks_spatial(masks, h, w).to(cuda_device), n_obj, inject_background)
    torch.cuda.synchronize()
    assert torch.equal(qd.float().cpu(), q_ref)
    assert torch.equal(kd.float().cpu(), k_ref)


@pytest.mark.parametrize("n_obj", [1, 2, 3])
@pytest.mark.parametrize("inject_background", [False, True])
def test_temporal_qk_blend(cuda_device, n_obj, inject_background):
    from mvoc_b200 import ops

    T, H, W, h, w, C = 8, 32, 32, 16, 16, 64
    nb = n_obj + 3
    masks = make_masks(n_obj, T, H, W, seed=10 + n_obj)
    torch.manual_seed(4)
    q = torch.randn(nb * h * w, T, C).bfloat16()
    k = torch.randn(nb * h * w, T, C).bfloat16()
    q_ref, k_ref = ops_ref.temporal_qk_inject_ref(q.float(), k.float(), masks, h, w, inject_background)
    qd, kd = q.to(cuda_device), k.to(cuda_device)
    ops.qk_blend_(qd, kd, _token_masks_temporal(masks, h, w).to(cuda_device), n_obj, inject_background)
    torch.cuda.synchronize()
    # fp32 lerp rounded once to bf16: equal to the rounded oracle up to 1 ulp ties
    assert rel_l2(qd, q_ref) <= 3e-3
    assert rel_l2(kd, k_ref) <= 3e-3
    assert torch.equal(qd.float().cpu()[: (n_obj + 1) * h * w], q.float()[: (n_obj + 1) * h * w])  # sources untouched


@pytest.mark.parametrize("n_obj", [1, 2, 3])
def test_feature_blend_bit_exact(cuda_device, n_obj):
    from mvoc_b200 import ops

    T, H, W, C = 8, 32, 32, 64
    nb = n_obj + 3
    masks = make_masks(n_obj, T, H, W, seed=20 + n_obj)
    torch.manual_seed(5)
    x = torch.randn(nb * T, C, H, W).bfloat16()
    ref = ops_ref.feature_inject_ref(x.float(), masks)
    m8 = torch.stack([mb[0, 0].reshape(T, H * W).to(torch.uint8) for _, mb in masks]).contiguous()
    xd = x.to(cuda_device)
    ops.feature_blend_(xd, m8.to(cuda_device), n_obj, T)
    torch.cuda.synchronize()
    assert torch.equal(xd.float().cpu(), ref)


# ------------------------------------------------------------------ GroupNorm
GN_CASES = [
    # N, C, H, W, frames_per_stat, silu, eps
    (16, 320, 32, 32, 1, True, 1e-5),     # slab path
    (16, 320, 32, 32, 1, False, 1e-6),    # transformer norm
    (8, 960, 64, 64, 1, True, 1e-5),      # slab too large for smem -> split path
    (16, 64, 16, 16, 8, True, 1e-5),      # temporal (5-D) statistics
    (16, 320, 32, 32, 8, False, 1e-6),
    (4, 1280, 8, 8, 1, True, 1e-5),
    (4, 64, 11, 20, 1, True, 1e-5),       # S % 8 != 0 -> scalar path
]


@pytest.mark.parametrize("N,C,H,W,fps,silu,eps", GN_CASES)
def test_groupnorm_silu_vs_oracle(cuda_device, N, C, H, W, fps, silu, eps):
    from mvoc_b200 import ops

    torch.manual_seed(C + H)
    x = (torch.randn(N, C, H, W) * 2.0 + 0.7).bfloat16()
    wgt = (1.0 + 0.2 * torch.randn(C)).bfloat16()
    bias = (0.1 * torch.randn(C)).bfloat16()
    ref = ops_ref.group_norm_ref(x.float(), wgt.float(), bias.float(), 32, eps, silu, fps)
    out = ops.groupnorm_silu(x.to(cuda_device), wgt.to(cuda_device), bias.to(cuda_device), 32, eps, silu, fps)
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    assert err <= 3e-3, f"rel L2 {err:.3e}"
    # in place
    xd = x.to(cuda_device)
    ops.groupnorm_silu(xd, wgt.to(cuda_device), bias.to(cuda_device), 32, eps, silu, fps, out=xd)
    torch.cuda.synchronize()
    assert torch.equal(xd, out)


# ------------------------------------------------------------------ latent kernels
@pytest.mark.parametrize("n_obj", [1, 2, 3])
@pytest.mark.parametrize("onf", [False, True])
def test_latent_composite_vs_oracle(cuda_device, n_obj, onf):
    from mvoc_b200 import ops

    T, h, w = 8, 32, 32
    masks = make_masks(n_obj, T, h, w, seed=30 + n_obj)
    torch.manual_seed(6)
    z = torch.randn(1, 4, T, h, w)
    bg = torch.randn(1, 4, T, h, w)
    objs = [torch.randn(1, 4, T, h, w) for _ in range(n_obj)]
    ratio = 0.3
    ref = ops_ref.latent_fusion_ref(z, bg, objs, [m for m, _ in masks], ratio, onf)
    zd, bgd = z.to(cuda_device).clone(), bg.to(cuda_device)
    objd = torch.cat(objs).to(cuda_device).contiguous()
    md = torch.stack([m[0, 0].reshape(-1) for m, _ in masks]).float().contiguous().to(cuda_device)
    unet_in = torch.empty(n_obj + 3, 4, T, h, w, dtype=torch.bfloat16, device=cuda_device)
    ops.latent_composite_(zd, bgd, objd, md, unet_in, ratio, True, onf)
    torch.cuda.synchronize()
    assert rel_l2(zd, ref) <= 1e-6
    src_in = torch.cat([bg, *objs]).bfloat16()
    assert torch.equal(unet_in[: n_obj + 1].cpu(), src_in)          # sources: exact casts
    zb = zd.cpu().bfloat16()                                          # composite slots: the kernel's own z
    assert torch.equal(unet_in[n_obj + 1].cpu(), zb[0]) and torch.equal(unet_in[n_obj + 2].cpu(), zb[0])
    # non-fusion step: z untouched, pure concat
    z2 = z.to(cuda_device).clone()
    ops.latent_composite_(z2, bgd, objd, None, unet_in, ratio, False)
    torch.cuda.synchronize()
    assert torch.equal(z2.cpu(), z)
    assert torch.equal(unet_in.cpu(), torch.cat([bg, *objs, z, z]).bfloat16())


def test_cfg_ddim_step_vs_oracle(cuda_device):
    from mvoc_b200 import ops

    torch.manual_seed(8)
    E = 4 * 8 * 32 * 32
    u = torch.randn(E).bfloat16()
    c = torch.randn(E).bfloat16()
    x = torch.randn(E)
    a_t, a_prev, g = 0.00078403, 0.00349756, 9.0
    ref = ops_ref.ddim_step_ref(ops_ref.cfg_ref(u.float(), c.float(), g), x, a_t, a_prev)
    xd = x.to(cuda_device).clone()
    ops.cfg_ddim_step_(u.to(cuda_device), c.to(cuda_device), xd, g, a_t, a_prev)
    torch.cuda.synchronize()
    assert rel_l2(xd, ref) <= 1e-5
    # known answer: x=1, v=0.5, 981 -> 961 => 0.983932
    one = torch.ones(8, device=cuda_device)
    half = torch.full((8,), 0.5, device=cuda_device)
    ops.cfg_ddim_step_(half, None, one, 1.0, a_t, a_prev)
    torch.cuda.synchronize()
    assert abs(float(one[0]) - 0.983932) < 2e-5
    # inverse step shares the algebra
    x3 = x.to(cuda_device).clone()
    ops.ddim_inverse_step_(u.to(cuda_device), None, x3, 1.0, a_prev, a_t)
    torch.cuda.synchronize()
    ref3 = ops_ref.ddim_step_ref(u.float(), x, a_prev, a_t)
    assert rel_l2(x3, ref3) <= 1e-5


# ------------------------------------------------------------------ channels-last kernels
GN_NHWC_CASES = [
    # N, C, H, W, frames_per_stat, silu, eps, with_add
    (16, 320, 32, 32, 1, True, 1e-5, False),
    (16, 320, 32, 32, 1, True, 1e-5, True),      # resnet norm2: GN(h + temb)
    (16, 320, 32, 32, 8, False, 1e-6, False),    # temporal transformer norm (5-D statistics)
    (8, 960, 64, 64, 1, True, 1e-5, False),
    (16, 64, 16, 16, 8, True, 1e-5, False),      # reduced UNet: 2 channels per group
    (4, 1280, 8, 8, 1, True, 1e-5, True),
    (6, 2560, 16, 16, 1, True, 1e-5, False),
    (4, 640, 11, 20, 2, True, 1e-5, False),      # non power-of-two spatial size
]


@pytest.mark.parametrize("N,C,H,W,fps,silu,eps,with_add", GN_NHWC_CASES)
def test_groupnorm_nhwc_vs_oracle(cuda_device, N, C, H, W, fps, silu, eps, with_add):
    from mvoc_b200 import ops

    torch.manual_seed(C + H + fps)
    x = (torch.randn(N, C, H, W) * 2.0 + 0.7).bfloat16()
    wgt = (1.0 + 0.2 * torch.randn(C)).bfloat16()
    bias = (0.1 * torch.randn(C)).bfloat16()
    add = (0.5 * torch.randn(N, C)).bfloat16() if with_add else None
    xin = x.float() + (add.float()[:, :, None, None] if with_add else 0.0)
    ref = ops_ref.group_norm_ref(xin, wgt.float(), bias.float(), 32, eps, silu, fps)
    x_cl = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    out = ops.groupnorm_nhwc(x_cl, wgt.to(cuda_device), bias.to(cuda_device), 32, eps, silu, fps,
                             add.to(cuda_device) if with_add else None)
    torch.cuda.synchronize()
    err = rel_l2(out.permute(0, 3, 1, 2), ref)
    assert err <= 3e-3, f"rel L2 {err:.3e}"
    inplace = x_cl.clone()
    ops.groupnorm_nhwc(inplace, wgt.to(cuda_device), bias.to(cuda_device), 32, eps, silu, fps,
                       add.to(cuda_device) if with_add else None, out=inplace)
    torch.cuda.synchronize()
    assert torch.equal(inplace, out)


def test_groupnorm_nhwc_sharded_statistics(cuda_device):
    """Pixel shards + gathered partial statistics == the un-sharded result (multi-GPU temporal GroupNorm)."""
    from mvoc_b200 import ops

    torch.manual_seed(1)
    N, S, C, P = 16, 256, 320, 4
    x = (torch.randn(N, S, C) * 1.5 + 0.3).bfloat16().to(cuda_device)
    w = (1.0 + 0.2 * torch.randn(C)).bfloat16().to(cuda_device)
    b = (0.1 * torch.randn(C)).bfloat16().to(cuda_device)
    full = ops.groupnorm_nhwc(x, w, b, 32, 1e-5, True, 8)
    shards = [x[:, i * S // P:(i + 1) * S // P].contiguous() for i in range(P)]
    from mvoc_b200 import _cabi

    lib = _cabi.load()
    parts = []
    chunks, _ = ops._gnh_geometry(S // P, C, 0)
    for sh in shards:
        part = torch.empty(N, 32, chunks, 2, device=cuda_device)
        _cabi.check(lib.mvoc_groupnorm_nhwc_stats(sh.data_ptr(), None, part.data_ptr(), N, S // P, C, 32, 0, None), "stats")
        parts.append(part)
    allp = torch.stack(parts)
    outs = [ops.groupnorm_nhwc(sh, w, b, 32, 1e-5, True, 8, gather=lambda p_: allp) for sh in shards]
    torch.cuda.synchronize()
    got = torch.cat(outs, dim=1)
    assert rel_l2(got, full) <= 2e-3


@pytest.mark.parametrize("M,F", [(4096, 1280), (1000, 256), (77, 5120)])
def test_geglu_vs_torch(cuda_device, M, F):
    from mvoc_b200 import ops

    torch.manual_seed(M)
    x = (torch.randn(M, 2 * F) * 1.5).bfloat16()
    a, g = x.float().chunk(2, dim=-1)
    ref = a * torch.nn.functional.gelu(g)
    out = ops.geglu(x.to(cuda_device))
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("B,T,S,H", [(5, 16, 256, 5), (3, 8, 100, 2), (2, 32, 64, 10), (5, 4, 64, 1)])
def test_temporal_attention_frames_vs_oracle(cuda_device, B, T, S, H):
    """Frame-major rows (b, t, pixel) read in place == the reference's [(b h w), T, C] attention."""
    from mvoc_b200 import ops

    torch.manual_seed(B * T + S)
    C = H * 64
    qkv = torch.randn(B * T * S, 3 * C).bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]

    def to_ref(x):  # (b t p) c -> (b p) t c
        return x.float().view(B, T, S, C).permute(0, 2, 1, 3).reshape(B * S, T, C)

    ref = ops_ref.sdpa_ref(to_ref(q), to_ref(k), to_ref(v), H)
    ref = ref.view(B, S, T, C).permute(0, 2, 1, 3).reshape(B * T * S, C)
    d = qkv.to(cuda_device)
    out = ops.temporal_attention_frames(d[:, :C], d[:, C:2 * C], d[:, 2 * C:], H, B, T, S)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 1e-2


def test_feature_blend_channels_last_bit_exact(cuda_device):
    """Hidden-state injection on channels-last rows (the Q/K select kernel with base = background)."""
    from mvoc_b200 import ops, pnp_utils

    n_obj, T, H, W, C = 2, 8, 32, 32, 64
    nb = n_obj + 3
    masks = make_masks(n_obj, T, H, W, seed=41)
    torch.manual_seed(9)
    x = torch.randn(nb * T, C, H, W).bfloat16()
    ref = ops_ref.feature_inject_ref(x.float(), masks)
    xd = x.permute(0, 2, 3, 1).contiguous().to(cuda_device)
    md = [(mf.to(cuda_device), mb.to(cuda_device)) for mf, mb in masks]
    ops.qk_blend_(xd, None, pnp_utils._MASKS.tokens(md, H, W, soft=False), n_obj, True)
    torch.cuda.synchronize()
    assert torch.equal(xd.permute(0, 3, 1, 2).float().cpu(), ref)


@pytest.mark.parametrize("M,C", [(4096, 320), (1000, 640), (333, 1280), (64, 64), (50, 512), (7, 2048)])
def test_layernorm_vs_torch(cuda_device, M, C):
    from mvoc_b200 import ops

    torch.manual_seed(M + C)
    x = (torch.randn(M, C) * 1.7 + 0.4).bfloat16()
    w = (1.0 + 0.2 * torch.randn(C)).bfloat16()
    b = (0.1 * torch.randn(C)).bfloat16()
    ref = torch.nn.functional.layer_norm(x.float(), (C,), w.float(), b.float(), 1e-5)
    out = ops.layernorm(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device), 1e-5)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 3e-3
    again = ops.layernorm(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device), 1e-5)
    assert torch.equal(out, again)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("N,H,W,C", [(80, 32, 32, 640), (5, 8, 8, 1280), (3, 11, 20, 64), (1, 1, 1, 8), (2, 7, 3, 328)])
def test_upsample_nearest2x_is_exact(cuda_device, N, H, W, C, dtype):
    """Upsample2D's F.interpolate(scale_factor=2, mode='nearest') on channels-last rows: a pure copy, bit exact."""
    from mvoc_b200 import ops

    torch.manual_seed(N + H + W + C)
    x = torch.randn(N, H, W, C).to(dtype).to(cuda_device)
    ref = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest")
    ref = ref.permute(0, 2, 3, 1).to(dtype)
    out = ops.upsample_nearest2x(x)
    torch.cuda.synchronize()
    assert out.shape == (N, 2 * H, 2 * W, C) and out.is_contiguous()
    assert torch.equal(out, ref)
    with pytest.raises(ValueError):
        ops.upsample_nearest2x(x[0])                     # not [N, H, W, C]


# ------------------------------------------------------------------ dense work on tcgen05 (csrc/gemm_tc.cu)
GEMM_VARIANTS = [0, 1]      # one CTA per tile / CTA pairs (cta_group::2)


@pytest.mark.parametrize("variant", GEMM_VARIANTS)
@pytest.mark.parametrize("M,K,N,mode", [(128, 64, 64, 0), (1000, 320, 320, 2), (4096, 320, 960, 0), (777, 640, 640, 2),
                                        (2048, 1280, 1280, 1), (11600, 1024, 640, 0), (300, 128, 192, 2),
                                        (5000, 320, 1280, 2)])
def test_linear_vs_torch(cuda_device, M, K, N, mode, variant):
    """mvoc_linear: x @ w^T (+ bias) (+ residual), single rounding of the fp32 accumulator."""
    from mvoc_b200 import ops

    torch.manual_seed(M + K + N)
    x = torch.randn(M, K).bfloat16()
    w = (torch.randn(N, K) * K ** -0.5).bfloat16()
    b = torch.randn(N).bfloat16() if mode >= 1 else None
    r = torch.randn(M, N).bfloat16() if mode >= 2 else None
    ref = x.float() @ w.float().t()
    if b is not None:
        ref = ref + b.float()
    if r is not None:
        ref = ref + r.float()
    d = lambda t: None if t is None else t.to(cuda_device)
    out = ops.linear(d(x), d(w), d(b), d(r), variant=variant)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("variant", GEMM_VARIANTS)
def test_linear_strided_views(cuda_device, variant):
    """x and out as column slices of wider buffers (row strides ldx / ldo), residual with its own stride."""
    from mvoc_b200 import ops

    torch.manual_seed(3)
    M, K, N = 640, 128, 192
    xb = torch.randn(M, K + 64, device=cuda_device).bfloat16()
    ob = torch.zeros(M, N + 128, device=cuda_device).bfloat16()
    rb = torch.randn(M, N + 64, device=cuda_device).bfloat16()
    w = (torch.randn(N, K, device=cuda_device) * K ** -0.5).bfloat16()
    x, out, r = xb[:, :K], ob[:, 64:64 + N], rb[:, :N]
    ops.linear(x, w, None, r, out=out, variant=variant)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t() + r.float()
    assert rel_l2(out, ref) <= 3e-3
    assert float(ob[:, :64].abs().max()) == 0.0 and float(ob[:, 64 + N:].abs().max()) == 0.0   # nothing outside


@pytest.mark.parametrize("variant", GEMM_VARIANTS)
@pytest.mark.parametrize("N,H,W,ci,co,ci2,res", [(2, 64, 64, 64, 64, 0, False), (4, 32, 32, 128, 128, 0, True),
                                                 (3, 16, 16, 64, 320, 0, True), (2, 11, 20, 64, 64, 0, False),
                                                 (5, 8, 8, 64, 320, 128, False), (5, 16, 16, 128, 640, 64, True),
                                                 (2, 64, 64, 320, 320, 0, True)])
def test_conv3x3_vs_torch(cuda_device, N, H, W, ci, co, ci2, res, variant):
    """mvoc_conv3x3_nhwc (+ fused 1x1 shortcut source, + residual) against torch conv2d in fp32."""
    import torch.nn.functional as F

    from mvoc_b200 import ops

    torch.manual_seed(N * H + ci + co)
    x = torch.randn(N, H, W, ci).bfloat16()
    w = (torch.randn(co, ci, 3, 3) * (9 * ci) ** -0.5).bfloat16()
    b = torch.randn(co).bfloat16()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float(), padding=1).permute(0, 2, 3, 1)
    x2 = w2 = r = None
    if ci2:
        x2 = torch.randn(N, H, W, ci2).bfloat16()
        w2 = (torch.randn(co, ci2) * ci2 ** -0.5).bfloat16()
        ref = ref + x2.float() @ w2.float().t()
    if res:
        r = torch.randn(N, H, W, co).bfloat16()
        ref = ref + r.float()
    d = lambda t: None if t is None else t.to(cuda_device)
    out = ops.conv3x3(d(x), ops.conv_taps(d(w)), d(b), d(r), d(x2), d(w2), variant=variant)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("variant", GEMM_VARIANTS)
@pytest.mark.parametrize("B,T,S,ci,co,res", [(2, 16, 256, 64, 64, False), (5, 16, 64, 128, 128, True),
                                             (3, 8, 100, 64, 192, True), (5, 32, 8, 64, 64, False),
                                             (1, 1, 256, 64, 64, False), (5, 2, 4096, 64, 64, True)])
def test_temporal_conv3_vs_torch(cuda_device, B, T, S, ci, co, res, variant):
    """mvoc_temporal_conv3 against torch Conv3d (3,1,1) with zero padding at the ends of each video."""
    import torch.nn.functional as F

    from mvoc_b200 import ops

    torch.manual_seed(B * T + S)
    x = torch.randn(B * T, S, ci).bfloat16()
    w = (torch.randn(co, ci, 3, 1, 1) * (3 * ci) ** -0.5).bfloat16()
    b = torch.randn(co).bfloat16()
    x5 = x.float().view(B, T, S, ci).permute(0, 3, 1, 2)[..., None]                    # [B, ci, T, S, 1]
    ref = F.conv3d(x5, w.float(), b.float(), padding=(1, 0, 0))[..., 0].permute(0, 2, 3, 1).reshape(B * T, S, co)
    r = None
    if res:
        r = torch.randn(B * T, S, co).bfloat16()
        ref = ref + r.float()
    d = lambda t: None if t is None else t.to(cuda_device)
    out = ops.temporal_conv3(d(x), ops.conv_taps(d(w)), d(b), B, T, d(r), variant=variant)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 3e-3


@pytest.mark.parametrize("variant", GEMM_VARIANTS)
@pytest.mark.parametrize("M,K,F", [(256, 64, 64), (1000, 320, 1280), (4096, 640, 2560), (300, 128, 192)])
def test_linear_geglu_vs_torch(cuda_device, M, K, F, variant):
    """GEGLU projection with the gate in the GEMM epilogue against Linear -> chunk -> value * gelu(gate) in fp32."""
    import torch.nn.functional as Fn

    from mvoc_b200 import ops

    torch.manual_seed(M + F)
    x = torch.randn(M, K).bfloat16()
    w = (torch.randn(2 * F, K) * K ** -0.5).bfloat16()
    b = torch.randn(2 * F).bfloat16()
    val, gate = (x.float() @ w.float().t() + b.float()).chunk(2, dim=-1)
    ref = val * Fn.gelu(gate)
    out = ops.linear_geglu(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device), variant=variant)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) <= 3e-3


def test_dense_fp16_storage(cuda_device):
    """fp16 activations / weights through the same kernels (the reference's GPU dtype)."""
    from mvoc_b200 import ops

    torch.manual_seed(5)
    x = torch.randn(512, 128).half()
    w = (torch.randn(256, 128) * 128 ** -0.5).half()
    b = torch.randn(256).half()
    out = ops.linear(x.to(cuda_device), w.to(cuda_device), b.to(cuda_device))
    torch.cuda.synchronize()
    assert out.dtype == torch.float16
    assert rel_l2(out, x.float() @ w.float().t() + b.float()) <= 1e-3
