"""TEST-ONLY emulation of the C-ABI kernel contracts in plain torch (fp32, CPU).

The product (`mvoc_b200/`) has no CPU path: `mvoc_b200.ops` raises on non-CUDA tensors.  To test the HOST
logic without a GPU — the channels-last UNet wiring, the hook layer, the step loops and the multi-GPU
partition — the tests in `tests/test_host_pipeline_cpu.py` monkeypatch `mvoc_b200.ops` with the functions
below, which restate each kernel's documented contract (include/mvoc_b200.h) with the oracle's ops.  Nothing
under `mvoc_b200/` imports this file.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from oracle import ops_ref


def _rows_ok(t, heads):
    """The layout the real wrappers accept: [.., heads*64] with a contiguous last dim and 16-byte-aligned rows."""
    assert t.shape[-1] == heads * 64 and t.stride(-1) == 1, (tuple(t.shape), t.stride())
    assert all(s % 8 == 0 for s in t.stride()[:-1]), t.stride()


def attention(q, k, v, heads, scale=None, out=None, variant=0):
    for t in (q, k, v):
        assert t.dim() == 3
        _rows_ok(t, heads)
    assert k.shape[0] == q.shape[0] and v.shape == k.shape
    o = ops_ref.sdpa_ref(q.contiguous(), k.contiguous(), v.contiguous(), heads)
    if out is not None:
        out.copy_(o)
        return out
    return o


def temporal_attention_frames(q, k, v, heads, B, T, S, scale=None, out=None):
    C = q.shape[-1]
    for t in (q, k, v):
        assert t.dim() == 2 and t.shape[0] == B * T * S and T <= 32
        _rows_ok(t, heads)

    def to_ref(x):  # rows (b, t, p) -> [(b p), t, c]
        return x.reshape(B, T, S, C).permute(0, 2, 1, 3).reshape(B * S, T, C)

    o = ops_ref.sdpa_ref(to_ref(q), to_ref(k), to_ref(v), heads)
    return o.view(B, S, T, C).permute(0, 2, 1, 3).reshape(B * T * S, C)


def temporal_attention(q, k, v, heads, scale=None, out=None):
    return ops_ref.sdpa_ref(q, k, v, heads)


def qk_blend_(q, k, mask, n_obj, inject_background):
    nb = n_obj + 3
    base = 0 if inject_background else n_obj + 2
    assert q.is_contiguous() and (k is None or k.is_contiguous()) and mask.is_contiguous()
    assert mask.dtype in (torch.uint8, torch.float32) and tuple(mask.shape) == (n_obj, q.numel() // (nb * q.shape[-1]))
    for x in (q, k):
        if x is None:
            continue
        C = x.shape[-1]
        v = x.view(nb, -1, C)
        acc = v[base].clone()
        for j in range(n_obj):
            m = mask[j].to(torch.float32)[:, None]
            if mask.dtype == torch.uint8:
                acc = torch.where(m != 0, v[j + 1], acc)
            else:
                acc = acc * (1 - m) + v[j + 1] * m
        v[n_obj + 1] = acc
        v[n_obj + 2] = acc


def attention_pair(q, k, v, heads, pair_batches, scale=None, out=None, variant=0):
    """out[b] = softmax(q[b] k[b]^T) v[b]; out[b + pair_batches] = the same softmax times v[b + pair_batches]."""
    B = q.shape[0]
    assert k.shape[0] == B and v.shape[0] == B + pair_batches and pair_batches >= 1
    o = torch.empty((B + pair_batches,) + tuple(q.shape[1:]), dtype=q.dtype)
    o[:B] = ops_ref.sdpa_ref(q.contiguous(), k.contiguous(), v[:B].contiguous(), heads)
    o[pair_batches:] = ops_ref.sdpa_ref(q.contiguous(), k.contiguous(), v[pair_batches:].contiguous(), heads)
    return o


def attention_inject_(q, k, v, mask, heads, n_obj, frames, inject_background, temporal, scale=None, out=None,
                      variant=0, share_p=None):
    """Contract of mvoc_attn_inject_fwd: blend (one copy with share_p) + attention of every branch; q, k, v are
    uniformly strided [(n_obj+3)*frames, pixels, C] views (column slices of a fused QKV buffer allowed)."""
    nb = n_obj + 3
    pixels, C = q.shape[1], q.shape[2]
    if share_p is None:
        share_p = not temporal
    for t in (q, k, v):
        assert tuple(t.shape) == (nb * frames, pixels, C) and t.stride(2) == 1 and t.stride(0) == pixels * t.stride(1)
        assert t.stride(1) % 8 == 0
    assert tuple(mask.shape) == (n_obj, frames * pixels) and mask.is_contiguous()
    assert not (share_p and temporal)
    base = 0 if inject_background else n_obj + 2
    blended = []
    for x in (q, k):
        vw = x.reshape(nb, frames * pixels, C)
        acc = vw[base].clone()
        for j in range(n_obj):
            m = mask[j].to(torch.float32)[:, None]
            acc = torch.where(m != 0, vw[j + 1], acc) if mask.dtype == torch.uint8 else acc * (1 - m) + vw[j + 1] * m
        blended.append(acc.view(frames, pixels, C))
    qq = torch.cat([q[:(n_obj + 1) * frames], blended[0], blended[0]], dim=0)
    kk = torch.cat([k[:(n_obj + 1) * frames], blended[1], blended[1]], dim=0)
    if temporal:
        return temporal_attention_frames(qq.reshape(-1, C), kk.reshape(-1, C), v.reshape(-1, C).contiguous(), heads, nb,
                                         frames, pixels).view(nb * frames, pixels, C)
    return ops_ref.sdpa_ref(qq, kk, v.contiguous(), heads)


def feature_blend_(x, mask, n_obj, frames):
    nb = n_obj + 3
    assert x.dim() == 4 and x.is_contiguous() and x.shape[0] == nb * frames
    assert mask.dtype == torch.uint8 and tuple(mask.shape) == (n_obj, frames, x.shape[2] * x.shape[3])
    v = x.view(nb, frames, x.shape[1], -1)                      # [slot, T, C, HW]
    acc = v[0].clone()
    for j in range(n_obj):
        m = mask[j].bool()[:, None, :]                          # [T, 1, HW]
        acc = torch.where(m, v[j + 1], acc)
    v[n_obj + 1] = acc
    v[n_obj + 2] = acc


def groupnorm_nhwc(x, weight, bias, groups, eps, silu, frames_per_stat=1, add=None, out=None, gather=None):
    """Same three phases as the kernels: per-(n, group) partial (mean, M2) of the local rows, optional gather of
    the partial sets of other ranks, Chan merge over sets and over the `frames_per_stat` frames, apply."""
    N, C = x.shape[0], x.shape[-1]
    assert x.is_contiguous() and C % groups == 0 and N % frames_per_stat == 0
    assert add is None or (tuple(add.shape) == (N, C) and add.is_contiguous() and add.dtype == x.dtype)
    cg = C // groups
    xs = x.reshape(N, -1, C).to(torch.float32)
    if add is not None:
        xs = xs + add.to(torch.float32)[:, None, :]
    S = xs.shape[1]
    xg = xs.view(N, S, groups, cg)
    mean = xg.mean(dim=(1, 3))
    m2 = ((xg - mean[:, None, :, None]) ** 2).sum(dim=(1, 3))
    partial = torch.stack([mean, m2], dim=-1)[:, :, None, :].contiguous()       # [N, G, 1, 2]
    sets = partial[None] if gather is None else gather(partial)                # [sets, N, G, 1, 2]
    cnt = float(S * cg)
    n_sets = sets.shape[0]
    pm = sets[..., 0, 0].reshape(n_sets, N // frames_per_stat, frames_per_stat, groups)
    pq = sets[..., 0, 1].reshape(n_sets, N // frames_per_stat, frames_per_stat, groups)
    total = cnt * n_sets * frames_per_stat
    gmean = pm.mean(dim=(0, 2))                                                  # equal counts per partial
    gm2 = pq.sum(dim=(0, 2)) + cnt * ((pm - gmean[None, :, None, :]) ** 2).sum(dim=(0, 2))
    rstd = torch.rsqrt(gm2 / total + eps)                                        # [N/frames, G]
    gmean = gmean.repeat_interleave(frames_per_stat, dim=0)[:, None, :, None]
    rstd = rstd.repeat_interleave(frames_per_stat, dim=0)[:, None, :, None]
    y = (xg - gmean) * rstd
    y = y.reshape(N, S, C) * weight.to(torch.float32) + bias.to(torch.float32)
    if silu:
        y = F.silu(y)
    y = y.to(x.dtype).view(x.shape)
    if out is not None:
        out.copy_(y)
        return out
    return y


def layernorm(x, weight, bias, eps, out=None):
    assert x.is_contiguous()
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def geglu(x, out=None):
    assert x.is_contiguous() and x.shape[-1] % 16 == 0
    a, g = x.chunk(2, dim=-1)
    return a * F.gelu(g)


def latent_composite_(z, bg, objs, mask, unet_in, ratio, do_fusion, obj_noise_fusion=False):
    n_obj = objs.shape[0]
    if do_fusion:
        zz = ratio * z + (1.0 - ratio) * bg.view_as(z)
        for j in range(n_obj):
            m = mask[j].view(1, 1, *z.shape[2:]).expand_as(z)
            obj = objs[j].view_as(z)
            fg = (zz * m) * ratio + (1 - ratio) * (obj * m) if obj_noise_fusion else obj * m
            zz = zz * (1.0 - m) + fg
        z.copy_(zz)
    if unet_in is not None:
        unet_in[0].copy_(bg.reshape(unet_in[0].shape))
        for j in range(n_obj):
            unet_in[j + 1].copy_(objs[j].reshape(unet_in[0].shape))
        unet_in[n_obj + 1].copy_(z.reshape(unet_in[0].shape))
        unet_in[n_obj + 2].copy_(z.reshape(unet_in[0].shape))


def _ddim(pu, pc, x, g, a_from, a_to):
    v = pu.to(torch.float32) if pc is None else pu.to(torch.float32) + g * (pc.to(torch.float32) - pu.to(torch.float32))
    x.copy_(ops_ref.ddim_step_ref(v.view_as(x), x, a_from, a_to))


def cfg_ddim_step_(pred_uncond, pred_cond, x, guidance, alpha_t, alpha_prev):
    _ddim(pred_uncond, pred_cond, x, guidance, alpha_t, alpha_prev)


def ddim_inverse_step_(pred_uncond, pred_cond, x, guidance, alpha_src, alpha_dst):
    _ddim(pred_uncond, pred_cond, x, guidance, alpha_src, alpha_dst)


# ---- dense work (csrc/gemm_tc.cu): contracts of mvoc_linear / mvoc_linear_geglu / mvoc_conv3x3_nhwc /
# mvoc_temporal_conv3.  `calls` counts what the host routed here (tests assert on it).
calls = {"linear": 0, "linear_res": 0, "geglu": 0, "conv": 0, "conv_res": 0, "conv_shortcut": 0, "tconv": 0,
         "tconv_res": 0}


def _aligned(*dims):
    assert all(d % 64 == 0 for d in dims), dims


def linear(x, weight, bias=None, residual=None, out=None, variant=None):
    n, k = weight.shape
    _aligned(n, k)
    assert weight.is_contiguous() and x.shape[-1] == k and x.stride(-1) == 1
    assert out is None or (out.data_ptr() != x.data_ptr() and (residual is None or out.data_ptr() != residual.data_ptr()))
    y = F.linear(x, weight, bias)
    calls["linear"] += 1
    if residual is not None:
        assert residual.shape == y.shape and residual.stride(-1) == 1
        calls["linear_res"] += 1
        y = y + residual
    if out is not None:
        out.copy_(y)
        return out
    return y


def linear_geglu(x, weight, bias=None, out=None, variant=None):
    f2, k = weight.shape
    _aligned(f2 // 2, k)
    assert x.is_contiguous() and weight.is_contiguous()
    calls["geglu"] += 1
    v, g = F.linear(x, weight, bias).chunk(2, dim=-1)
    return v * F.gelu(g)


def conv3x3(x, w_taps, bias=None, residual=None, x2=None, w2=None, out=None, variant=None):
    co, ci = w_taps.shape[1], w_taps.shape[2]
    _aligned(co, ci)
    assert x.dim() == 4 and x.is_contiguous() and w_taps.is_contiguous() and w_taps.shape[0] == 9
    w = w_taps.view(3, 3, co, ci).permute(2, 3, 0, 1)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, bias, padding=1).permute(0, 2, 3, 1).contiguous()
    calls["conv"] += 1
    if x2 is not None:
        c2 = x2.shape[-1]
        _aligned(c2)
        assert x2.is_contiguous() and tuple(x2.shape[:3]) == tuple(x.shape[:3]) and tuple(w2.shape) == (co, c2)
        calls["conv_shortcut"] += 1
        y = y + F.linear(x2, w2)
    if residual is not None:
        assert residual.shape == y.shape and residual.is_contiguous()
        calls["conv_res"] += 1
        y = y + residual
    return y


def temporal_conv3(x, w_taps, bias, videos, frames, residual=None, out=None, variant=None):
    co, ci = w_taps.shape[1], w_taps.shape[2]
    _aligned(co, ci)
    assert x.is_contiguous() and x.shape[0] == videos * frames and w_taps.shape[0] == 3 and w_taps.is_contiguous()
    xs = x.reshape(videos, frames, -1, ci)                                     # [b, t, s, ci]
    y = torch.zeros(videos, frames, xs.shape[2], co, dtype=x.dtype)
    for tap in range(3):
        sh = tap - 1                                                           # frame offset of this tap
        lo, hi = max(0, -sh), min(frames, frames - sh)
        if hi > lo:
            y[:, lo:hi] += F.linear(xs[:, lo + sh:hi + sh], w_taps[tap])
    if bias is not None:
        y = y + bias
    y = y.reshape(x.shape[:-1] + (co,))
    calls["tconv"] += 1
    if residual is not None:
        assert residual.shape == y.shape and residual.is_contiguous()
        calls["tconv_res"] += 1
        y = y + residual
    return y


def install(monkeypatch_or_module=None, dense: bool = True):
    """Replace the kernel wrappers of mvoc_b200.ops with the emulations (tests only).  With `dense` the host's
    routing predicate for the tcgen05 GEMM family is opened for CPU tensors too, so the weight re-layouts and
    epilogue fusions of the product path (tap-major filters, fused QKV, residual / shortcut / GEGLU in the
    epilogue) are what runs; without it the cuDNN / cuBLAS branch (MVOC_DENSE=lib) runs."""
    from mvoc_b200 import ops, unet3d

    names = ["attention", "attention_pair", "attention_inject_", "temporal_attention_frames", "temporal_attention",
             "qk_blend_", "feature_blend_",
             "layernorm", "geglu", "latent_composite_", "cfg_ddim_step_", "ddim_inverse_step_",
             "linear", "linear_geglu", "conv3x3", "temporal_conv3"]
    saved = {n: getattr(ops, n) for n in names}
    for n in names:
        setattr(ops, n, globals()[n])
    # ops.groupnorm_nhwc itself (the slab-slicing wrapper) stays real; the kernel-calling piece is emulated
    saved["_groupnorm_nhwc_slab"] = ops._groupnorm_nhwc_slab
    ops._groupnorm_nhwc_slab = groupnorm_nhwc
    saved["_tc_ok"] = unet3d._tc_ok
    unet3d._tc_ok = (lambda x, k, n: k % 64 == 0 and n % 64 == 0) if dense else (lambda x, k, n: False)
    for key in calls:
        calls[key] = 0
    return saved


def uninstall(saved):
    from mvoc_b200 import ops, unet3d

    for n, f in saved.items():
        if n == "_tc_ok":
            unet3d._tc_ok = f
        else:
            setattr(ops, n, f)
