"""Size-independent properties at BASELINE config-2 FULL sizes (80 frames x 64x64 latents, C = 320, 5 heads),
where an fp32 CPU oracle run would take minutes: softmax rows sum to one, key-permutation invariance,
blend idempotence, unit statistics after normalisation, DDIM step round trip.  All through the C-ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu

B, H, N, C, T, NOBJ = 80, 5, 4096, 320, 16, 2     # l0 of config 2: 5 branches x 16 frames, 64x64 tokens


def test_attention_full_size_constant_v_and_key_permutation(cuda_device):
    from mvoc_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(0)
    q = torch.randn(B, N, C, device=cuda_device, generator=g).bfloat16()
    k = torch.randn(B, N, C, device=cuda_device, generator=g).bfloat16()
    # (1) V constant along the key axis => O equals that constant for every query (rows of softmax sum to 1)
    vconst = torch.randn(B, 1, C, device=cuda_device, generator=g).bfloat16().expand(B, N, C).contiguous()
    o = ops.attention(q, k, vconst, H)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all()
    err = (o.float() - vconst.float()).abs().max().item()
    assert err <= 2e-2 * vconst.float().abs().max().item() + 1e-3, err
    # (2) permuting keys and values together leaves the output unchanged (up to accumulation order)
    v = torch.randn(B, N, C, device=cuda_device, generator=g).bfloat16()
    o1 = ops.attention(q, k, v, H)
    perm = torch.randperm(N, device=cuda_device, generator=g)
    o2 = ops.attention(q, k[:, perm].contiguous(), v[:, perm].contiguous(), H)
    torch.cuda.synchronize()
    rel = ((o1.float() - o2.float()).norm() / o1.float().norm()).item()
    assert rel <= 1e-2, rel
    # (3) the five exp2 / MMA variants agree on the full-size problem
    o3 = ops.attention(q, k, v, H, variant=2)
    torch.cuda.synchronize()
    assert ((o1.float() - o3.float()).norm() / o3.float().norm()).item() <= 5e-3


def test_temporal_attention_full_size_constant_v(cuda_device):
    from mvoc_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(1)
    S, Bv = 4096, NOBJ + 3
    rows = Bv * T * S
    q = torch.randn(rows, C, device=cuda_device, generator=g).bfloat16()
    k = torch.randn(rows, C, device=cuda_device, generator=g).bfloat16()
    # value constant over the frames of each (video, pixel): rows are (video, frame, pixel)
    base = torch.randn(Bv, 1, S, C, device=cuda_device, generator=g).bfloat16()
    v = base.expand(Bv, T, S, C).reshape(rows, C).contiguous()
    o = ops.temporal_attention_frames(q, k, v, H, Bv, T, S)
    torch.cuda.synchronize()
    err = (o.float() - v.float()).abs().max().item()
    assert err <= 2e-2 * v.float().abs().max().item() + 1e-3, err


def test_qk_blend_full_size_idempotent_and_zero_mask(cuda_device):
    from mvoc_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(2)
    nb = NOBJ + 3
    q = torch.randn(nb * T, N, C, device=cuda_device, generator=g).bfloat16()
    k = torch.randn(nb * T, N, C, device=cuda_device, generator=g).bfloat16()
    mask = (torch.rand(NOBJ, T * N, device=cuda_device, generator=g) < 0.08).to(torch.uint8)
    src_q = q[: (NOBJ + 1) * T].clone()
    ops.qk_blend_(q, k, mask, NOBJ, False)
    once_q, once_k = q.clone(), k.clone()
    ops.qk_blend_(q, k, mask, NOBJ, False)                   # a second application changes nothing
    torch.cuda.synchronize()
    assert torch.equal(q, once_q) and torch.equal(k, once_k)
    assert torch.equal(q[: (NOBJ + 1) * T], src_q)           # sources are never written
    u, c = q[(NOBJ + 1) * T:(NOBJ + 2) * T], q[(NOBJ + 2) * T:]
    assert torch.equal(u, c)                                 # uncond == cond after injection (pnp_utils.py:664-668)
    # where object j's mask is set (and no later object's), the composite rows are object j's rows
    m = mask.view(NOBJ, T, N).bool()
    only1 = m[0] & ~m[1]
    assert torch.equal(c[only1], q[T:2 * T][only1])
    assert torch.equal(c[m[1]], q[2 * T:3 * T][m[1]])
    # zero masks + cond base: nothing but the uncond <- cond copy happens
    q2 = torch.randn(nb * T, 256, C, device=cuda_device, generator=g).bfloat16()
    ref = q2.clone()
    ops.qk_blend_(q2, None, torch.zeros(NOBJ, T * 256, dtype=torch.uint8, device=cuda_device), NOBJ, False)
    torch.cuda.synchronize()
    assert torch.equal(q2[(NOBJ + 2) * T:], ref[(NOBJ + 2) * T:])
    assert torch.equal(q2[(NOBJ + 1) * T:(NOBJ + 2) * T], ref[(NOBJ + 2) * T:])
    assert torch.equal(q2[: (NOBJ + 1) * T], ref[: (NOBJ + 1) * T])


@pytest.mark.parametrize("frames_per_stat", [1, 16])
def test_groupnorm_full_size_unit_statistics(cuda_device, frames_per_stat):
    from mvoc_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(3)
    x = (torch.randn(B, 64, 64, C, device=cuda_device, generator=g) * 3.0 + 1.5).bfloat16()
    w = torch.ones(C, device=cuda_device).bfloat16()
    b = torch.zeros(C, device=cuda_device).bfloat16()
    y = ops.groupnorm_nhwc(x, w, b, 32, 1e-5, False, frames_per_stat)
    torch.cuda.synchronize()
    yy = y.float().view(B // frames_per_stat, frames_per_stat * 4096, 32, C // 32)
    mean = yy.mean(dim=(1, 3))
    var = yy.var(dim=(1, 3), unbiased=False)
    assert mean.abs().max().item() <= 5e-3
    assert (var - 1.0).abs().max().item() <= 1e-2
    # per-channel affine + shift invariance: GN(x + c) == GN(x) for a per-(frame, channel) constant only if the
    # constant is uniform inside a group; use the fused `add` path with a per-group constant
    add = torch.randn(B, 32, 1, device=cuda_device, generator=g).expand(B, 32, C // 32).reshape(B, C).bfloat16().contiguous()
    if frames_per_stat == 1:
        y2 = ops.groupnorm_nhwc(x, w, b, 32, 1e-5, False, 1, add=add)
        torch.cuda.synchronize()
        assert ((y2.float() - y.float()).norm() / y.float().norm()).item() <= 1e-2


def test_layernorm_and_geglu_full_size(cuda_device):
    from mvoc_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(4)
    M = B * N
    x = (torch.randn(M, C, device=cuda_device, generator=g) * 2.0 - 0.7).bfloat16()
    y = ops.layernorm(x, torch.ones(C, device=cuda_device).bfloat16(), torch.zeros(C, device=cuda_device).bfloat16(), 1e-5)
    torch.cuda.synchronize()
    yf = y.float()
    assert yf.mean(dim=1).abs().max().item() <= 1e-2
    assert (yf.var(dim=1, unbiased=False) - 1.0).abs().max().item() <= 2e-2
    del yf, y
    # GEGLU: gate = 0 -> 0; large positive gate -> x * gate
    F = 4 * C
    a = torch.randn(M // 8, F, device=cuda_device, generator=g).bfloat16()
    z = ops.geglu(torch.cat([a, torch.zeros_like(a)], dim=1).contiguous())
    big = torch.full_like(a, 8.0)
    z2 = ops.geglu(torch.cat([a, big], dim=1).contiguous())
    torch.cuda.synchronize()
    assert z.float().abs().max().item() == 0.0
    assert ((z2.float() - 8.0 * a.float()).norm() / (8.0 * a.float()).norm()).item() <= 5e-3


def test_ddim_round_trip_and_composite_identity_full_size(cuda_device):
    from mvoc_b200 import ops
    from mvoc_b200.scheduler import DDIMSchedule

    g = torch.Generator(device=cuda_device).manual_seed(5)
    E = 4 * T * 64 * 64
    s = DDIMSchedule(50)
    a_t, a_prev = s.step_alphas(501)
    x = torch.randn(E, device=cuda_device, generator=g)
    v = torch.randn(E, device=cuda_device, generator=g)
    x0 = (a_t ** 0.5) * x - ((1 - a_t) ** 0.5) * v
    eps = (a_t ** 0.5) * v + ((1 - a_t) ** 0.5) * x
    x1 = x.clone()
    ops.cfg_ddim_step_(v, None, x1, 1.0, a_t, a_prev)                  # level t -> t_prev
    v_prev = (a_prev ** 0.5) * eps - ((1 - a_prev) ** 0.5) * x0          # the same (x0, eps) seen from t_prev
    ops.ddim_inverse_step_(v_prev, None, x1, 1.0, a_prev, a_t)          # and back
    torch.cuda.synchronize()
    assert ((x1 - x).norm() / x.norm()).item() <= 1e-5
    # guidance 1.0 with pred_cond == pred_uncond is the unguided step
    x2, x3 = x.clone(), x.clone()
    ops.cfg_ddim_step_(v, v.clone(), x2, 9.0, a_t, a_prev)
    ops.cfg_ddim_step_(v, None, x3, 1.0, a_t, a_prev)
    torch.cuda.synchronize()
    assert torch.allclose(x2, x3, atol=1e-6)
    # latent composite with ratio 1 and empty masks leaves z alone and fills the five UNet-input slots
    z = torch.randn(1, 4, T, 64, 64, device=cuda_device, generator=g)
    bg = torch.randn_like(z)
    objs = torch.randn(NOBJ, 4, T, 64, 64, device=cuda_device, generator=g)
    masks = torch.zeros(NOBJ, T * 4096, device=cuda_device)
    unet_in = torch.empty(NOBJ + 3, 4, T, 64, 64, dtype=torch.bfloat16, device=cuda_device)
    z0 = z.clone()
    ops.latent_composite_(z, bg, objs, masks, unet_in, 1.0, True)
    torch.cuda.synchronize()
    assert torch.equal(z, z0)
    assert torch.equal(unet_in[0], bg[0].bfloat16()) and torch.equal(unet_in[1:3], objs.bfloat16())
    assert torch.equal(unet_in[3], z0[0].bfloat16()) and torch.equal(unet_in[4], z0[0].bfloat16())


@pytest.mark.parametrize("temporal,share_p,fused", [(False, False, False), (False, True, False), (False, True, True),
                                                    (True, False, False), (True, False, True)])
def test_attn_inject_entry_equals_blend_then_attention(cuda_device, temporal, share_p, fused):
    """mvoc_attn_inject_fwd (one C-ABI call) against mvoc_qk_blend followed by the attention entry point and against
    the fp32 oracle ops.  share_p: the uncond / cond pair runs through the one-softmax pair kernel (64-key blocks),
    so it agrees with the two-softmax path to rounding, not bit for bit.  fused: q, k, v are the column slices of
    ONE [rows, 3C] buffer, as the product's fused QKV projection hands them over."""
    from mvoc_b200 import ops
    from oracle import ops_ref
    from tests import cpu_ops_emulation as emu

    g = torch.Generator(device=cuda_device).manual_seed(6)
    n_obj, frames, pixels, heads = 2, 4, 256, 2
    nb, C = n_obj + 3, heads * 64
    qkv = torch.randn(nb * frames, pixels, 3 * C, device=cuda_device, generator=g).bfloat16()
    if fused:
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q, k, v = (qkv[..., i * C:(i + 1) * C].contiguous() for i in range(3))
    m = torch.rand(n_obj, frames * pixels, device=cuda_device, generator=g)
    mask = m.clamp(0, 1).contiguous() if temporal else (m < 0.1).to(torch.uint8)
    # two validated calls on contiguous copies
    q2, k2, v2 = q.contiguous().clone(), k.contiguous().clone(), v.contiguous()
    ops.qk_blend_(q2, k2, mask, n_obj, False)
    if temporal:
        ref = ops.temporal_attention_frames(q2.view(-1, C), k2.view(-1, C), v2.view(-1, C), heads, nb, frames,
                                            pixels).view_as(q2)
    else:
        ref = ops.attention(q2, k2, v2, heads)
    # the single entry point (modifies q, k in place: work on a copy of the fused buffer)
    qkv1 = qkv.clone()
    if fused:
        q1, k1, v1 = qkv1[..., :C], qkv1[..., C:2 * C], qkv1[..., 2 * C:]
    else:
        q1, k1, v1 = q.clone(), k.clone(), v
    out = ops.attention_inject_(q1, k1, v1, mask, heads, n_obj, frames, False, temporal, share_p=share_p)
    torch.cuda.synchronize()
    u = slice((n_obj + 1) * frames, (n_obj + 2) * frames)
    assert torch.equal(q1[u], q2[u]) and torch.equal(k1[u], k2[u])           # the blended copy the kernels read
    if not share_p:
        assert torch.equal(q1, q2) and torch.equal(k1, k2)
        assert torch.equal(out, ref)
    else:
        src = slice(0, (n_obj + 1) * frames)
        assert torch.equal(out[src], ref[src])                                # source branches: the same kernel
        rel_pair = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
        assert rel_pair <= 5e-3, rel_pair
    # fp32 oracle on the same bf16-rounded inputs
    qc, kc, vc = q.float().cpu().contiguous(), k.float().cpu().contiguous(), v.float().cpu().contiguous()
    emu.qk_blend_(qc, kc, mask.cpu(), n_obj, False)
    if temporal:
        want = emu.temporal_attention_frames(qc.view(-1, C), kc.view(-1, C), vc.view(-1, C), heads, nb, frames,
                                             pixels).view_as(qc)
    else:
        want = ops_ref.sdpa_ref(qc, kc, vc, heads)
    rel = ((out.float().cpu() - want).norm() / want.norm()).item()
    assert rel <= 1e-2, rel
