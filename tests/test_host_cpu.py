"""CPU-side tests: C-ABI surface, host logic (scheduler, config, injection sites), oracle unit pins."""
import ctypes
import json
import os
import re
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C-ABI
def test_library_exports_every_declared_symbol():
    """The .so loads (no GPU needed) and exports exactly what include/mvoc_b200.h declares."""
    from mvoc_b200 import _cabi

    hdr = open(os.path.join(ROOT, "include", "mvoc_b200.h")).read()
    declared = set(re.findall(r"\b(mvoc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_cabi.SIGNATURES), (declared ^ set(_cabi.SIGNATURES))
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert "sm_100a" in _cabi.version()
    assert lib.mvoc_groupnorm_workspace_bytes(80, 32) >= 80 * 32 * 8


def test_argument_errors_are_reported_without_a_gpu():
    """Validation happens before any CUDA call: bad arguments return an error code + message."""
    from mvoc_b200 import _cabi

    lib = _cabi.load()
    rc = lib.mvoc_attn_fwd(None, None, None, None, 1, 1, 128, 128, 64, *([0] * 12), 0.125, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.mvoc_last_error()
    rc = lib.mvoc_attn_fwd(16, 16, 16, 16, 1, 1, 128, 128, 48, *([0] * 12), 0.125, 0, 0, None)
    assert rc == -2 and b"head_dim" in lib.mvoc_last_error()
    rc = lib.mvoc_qk_blend(16, 16, 0, 10, 64, 16, 0, 2, 0, None)
    assert rc == -1 and b"n_obj" in lib.mvoc_last_error()
    rc = lib.mvoc_cfg_ddim_step(16, 16, 16, 7, 1.0, 0.5, 0.6, 0, 2, None)
    assert rc == -2 and b"multiple of 8" in lib.mvoc_last_error()


def test_product_ops_refuse_cpu_tensors():
    """No CPU fallback: the product raises instead of computing with torch."""
    from mvoc_b200 import ops

    q = torch.randn(1, 128, 64).bfloat16()
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.attention(q, q, q, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.groupnorm_silu(torch.randn(2, 32, 4, 4).bfloat16(), torch.ones(32).bfloat16(), torch.zeros(32).bfloat16(),
                           32, 1e-5, True)
    from mvoc_b200.unet3d import I2VGenXLUNet, UNetConfig

    m = I2VGenXLUNet(UNetConfig.reduced())
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 4, 2, 8, 8), 1, torch.tensor([8]), torch.randn(1, 4, 2, 8, 8), torch.randn(1, 1, 1024),
          torch.randn(1, 77, 1024))


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mvoc_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"


# ------------------------------------------------------------------ scheduler
KNOWN_ALPHAS = {0: 0.999959, 1: 0.999913, 21: 0.997971, 501: 0.490706, 921: 0.0146935, 961: 0.00349756,
                981: 0.00078403, 999: 0.0}


def test_schedule_known_answers():
    """SURVEY App. A.6 known answers of the i2vgen-xl scheduler config."""
    from mvoc_b200.scheduler import DDIMSchedule
    from oracle import scheduler as osched

    s = DDIMSchedule(50)
    for t, a in KNOWN_ALPHAS.items():
        assert abs(s.alpha(t) - a) < 2e-6 * max(1.0, a / 1e-3), (t, s.alpha(t), a)
    assert s.timesteps[:4] == [981, 961, 941, 921] and s.timesteps[-1] == 1 and len(s.timesteps) == 50
    assert (s.alphas_cumprod == osched.alphas_cumprod().numpy()).all()
    inv = DDIMSchedule(500, inverse=True)
    assert inv.timesteps[:3] == [1, 3, 5] and inv.timesteps[-1] == 999
    assert set(s.timesteps) <= set(inv.timesteps)       # why inversion uses 500 steps
    assert inv.step_alphas(1) == (1.0, s.alpha(1))      # set_alpha_to_one below level 0
    o = osched.DDIMScheduler()
    o.set_timesteps(50)
    x = float(o.step(torch.tensor(0.5), 981, torch.tensor(1.0)))
    assert abs(x - 0.983932) < 2e-6                     # x=1, v=0.5, 981 -> 961
    # inverse then forward over the same pair of levels is the identity for a fixed prediction
    oi = osched.DDIMInverseScheduler()
    oi.set_timesteps(50)
    v = torch.tensor(0.3)
    x1 = oi.step(v, 21, torch.tensor(0.7))              # level 1 -> 21
    # map v (defined at level 1) to level 21 so that x0/eps are preserved, then step back
    a1, a21 = o.alphas_cumprod[1], o.alphas_cumprod[21]
    x0 = a1.sqrt() * 0.7 - (1 - a1).sqrt() * v
    eps = a1.sqrt() * v + (1 - a1).sqrt() * 0.7
    v21 = a21.sqrt() * eps - (1 - a21).sqrt() * x0
    back = o.step(v21, 21, x1)
    assert abs(float(back) - 0.7) < 1e-5


def test_init_pnp_schedules_follow_the_full_grid():
    """composite.py:39-52: k = int(n_steps * frac); schedule = timesteps[:k] of the FULL grid."""
    from mvoc_b200.pipeline import I2VGenXLPipeline, init_pnp
    from mvoc_b200.scheduler import DDIMSchedule
    from mvoc_b200.unet3d import I2VGenXLUNet, UNetConfig

    pipe = I2VGenXLPipeline(I2VGenXLUNet(UNetConfig.reduced()), "cpu")
    cfg = SimpleNamespace(n_steps=50, pnp_f_t=0.1, pnp_spatial_attn_t=0.5, pnp_temp_attn_t=0.0, inject_background=False)
    sch = init_pnp(pipe, DDIMSchedule(50), cfg)
    assert sch["conv"] == [981, 961, 941, 921, 901]
    assert len(sch["spatial"]) == 25 and sch["spatial"][-1] == 501
    assert sch["temporal"] == []          # an empty slice never fires
    from mvoc_b200 import pnp_utils

    blk = pipe.unet.up_blocks[-1]
    p = blk.attentions[1].transformer_blocks[0].attn1.processor
    assert type(p).__name__ == "ModifiedSpaAttnProcessor" and p.injection_schedule == sch["spatial"]
    # layer 0 of the lowest-resolution cross-attention up block keeps the stock processor (pnp_utils.py:706-707)
    assert type(blk.attentions[0].transformer_blocks[0].attn1.processor).__name__ == "AttnProcessor2_0"
    pnp_utils.register_time_all(pipe, 981, [])
    assert p.t == 981 and blk.resnets[0].t == 981 and pipe.unet.conv_out.t == 981
    p.t = 501
    assert pnp_utils._fires(p)
    p.t = 481
    assert not pnp_utils._fires(p)
    p.t = 1000
    assert pnp_utils._fires(p)            # `or self.t == 1000`


def test_full_unet_matches_published_parameter_count():
    from mvoc_b200.unet3d import I2VGenXLUNet, UNetConfig
    from oracle.unet import I2VGenXLUNet as OUNet, UNetConfig as OCfg

    with torch.device("meta"):
        p, o = I2VGenXLUNet(UNetConfig.full()), OUNet(OCfg.full())
    assert sum(x.numel() for x in p.parameters()) == 1_420_469_224
    assert sum(x.numel() for x in o.parameters()) == 1_420_469_224
    assert list(p.state_dict().keys()) == list(o.state_dict().keys())
    assert p.state_dict()["up_blocks.3.temp_convs.2.conv4.3.weight"].shape == (320, 320, 3, 1, 1)


# ------------------------------------------------------------------ config
def test_template_merge_and_interpolation(tmp_path):
    """OmegaConf subset used by the reference configs: deep merge, ${a} and ${a.b} interpolation, `active`."""
    from mvoc_b200 import config as cfg

    (tmp_path / "template.yaml").write_text(
        'seed: 6\ndata_dir: ".."\nmodel_name: "i2vgen-xl"\nvideo_name: "ReplaceMe"\n'
        'output_dir: "${data_dir}/Results/${model_name}/${video_name}/"\n'
        'n_steps: 50\npnp_f_t: 0.2\nfusion_step: [0, 3]\n'
        'frameinit_kwargs:\n  enable: true\n  filter_params:\n    method: "gaussian"\n    d_s: 0.25\n'
        'alias: "${frameinit_kwargs.filter_params.d_s}"\n')
    (tmp_path / "group.json").write_text(json.dumps([
        {"active": True, "video_name": "boat_surf", "pnp_f_t": 0.1, "fusion_step": [0, 1],
         "frameinit_kwargs": {"filter_params": {"d_s": 0.5}}},
        {"active": False, "video_name": "skipped"}]))
    out = list(cfg.iter_configs(str(tmp_path / "template.yaml"), str(tmp_path / "group.json")))
    assert len(out) == 1
    c = out[0]
    assert c.output_dir == "../Results/i2vgen-xl/boat_surf/"
    assert c.pnp_f_t == 0.1 and c.fusion_step == [0, 1] and c.n_steps == 50
    assert c.frameinit_kwargs.enable is True and c.frameinit_kwargs.filter_params.method == "gaussian"
    assert c.alias == 0.5                  # whole-value reference keeps the type


# ------------------------------------------------------------------ oracle unit pins
def _literal_two_object_spatial(query, key, masks, height, width, inject_background):
    """Literal transcription of pnp_utils.py:628-672 INCLUDING the hard-coded 5 (two objects only)."""
    import torch.nn.functional as F
    from einops import rearrange

    batch_size = query.shape[0]
    chunk_size = batch_size // 5
    query = rearrange(query.clone(), "b (h w) c -> b h w c", h=height)
    key = rearrange(key.clone(), "b (h w) c -> b h w c", h=height)
    if inject_background:
        q_inject, k_inject = query[:chunk_size], key[:chunk_size]
    else:
        q_inject, k_inject = query[4 * chunk_size:], key[4 * chunk_size:]
    for j, obj_mask_tensor in enumerate(masks):
        obj_q = query[chunk_size * (j + 1):chunk_size * (j + 2)]
        obj_k = key[chunk_size * (j + 1):chunk_size * (j + 2)]
        m = obj_mask_tensor[1].to(torch.float32)
        m = rearrange(m, "a b l h w -> (a b) l h w")
        m = F.interpolate(m, size=(height, width), mode="nearest")
        m = m[0]
        m = m.unsqueeze(-1).repeat(1, 1, 1, query.shape[-1])
        q_inject = q_inject * (1 - m) + obj_q * m
        k_inject = k_inject * (1 - m) + obj_k * m
    query[3 * chunk_size: 4 * chunk_size] = q_inject
    key[3 * chunk_size: 4 * chunk_size] = k_inject
    query[4 * chunk_size:] = q_inject
    key[4 * chunk_size:] = k_inject
    return rearrange(query, "b h w c -> b (h w) c"), rearrange(key, "b h w c -> b (h w) c")


@pytest.mark.parametrize("inject_background", [False, True])
def test_generalised_chunking_equals_the_hard_coded_five(inject_background):
    from mvoc_b200.synthetic import make_masks
    from oracle import ops_ref

    T, H, W, h, w, C = 4, 16, 16, 8, 8, 32
    masks = make_masks(2, T, H, W, seed=3)
    torch.manual_seed(0)
    q, k = torch.randn(5 * T, h * w, C), torch.randn(5 * T, h * w, C)
    a = ops_ref.spatial_qk_inject_ref(q, k, masks, h, w, inject_background)
    b = _literal_two_object_spatial(q, k, masks, h, w, inject_background)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0][: 3 * T], q[: 3 * T])              # sources untouched
    assert torch.equal(a[0][3 * T: 4 * T], a[0][4 * T:])        # uncond == cond after injection


def test_nearest_resize_rule():
    """F.interpolate(mode='nearest') picks src = floor(dst * in / out), also for non-integer ratios."""
    from oracle import ops_ref

    m = torch.arange(2 * 32 * 32, dtype=torch.float32).view(2, 32, 32)
    for (h, w) in [(16, 16), (8, 8), (11, 20)]:
        r = ops_ref.nearest_mask(m, h, w)
        iy = torch.floor(torch.arange(h) * (32 / h)).long()
        ix = torch.floor(torch.arange(w) * (32 / w)).long()
        assert torch.equal(r, m[:, iy][:, :, ix])


def test_later_object_wins_on_overlap():
    from oracle import ops_ref

    T, H, W = 2, 8, 8
    full = torch.ones(1, 4, T, H, W, dtype=torch.bool)
    masks = [(full.float(), full), (full.float(), full)]
    x = torch.arange(5 * T, dtype=torch.float32).view(5 * T, 1, 1, 1).expand(5 * T, 3, H, W).contiguous()
    y = ops_ref.feature_inject_ref(x, masks)
    assert torch.equal(y[3 * T:4 * T], x[2 * T:3 * T]) and torch.equal(y[4 * T:], x[2 * T:3 * T])


def test_latent_fusion_keeps_background_outside_masks():
    from mvoc_b200.synthetic import make_masks
    from oracle import ops_ref

    masks = make_masks(2, 4, 16, 16, seed=1)
    z, bg = torch.randn(1, 4, 4, 16, 16), torch.randn(1, 4, 4, 16, 16)
    objs = [torch.randn(1, 4, 4, 16, 16) for _ in range(2)]
    out = ops_ref.latent_fusion_ref(z, bg, objs, [m for m, _ in masks], ratio=0.0)
    outside = (masks[0][0] == 0) & (masks[1][0] == 0)
    assert torch.equal(out[outside], bg[outside])              # r = 0: pure background outside every mask
    inside1 = masks[1][0] == 1
    assert torch.allclose(out[inside1], objs[1][inside1])
