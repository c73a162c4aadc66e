"""Host logic of the product without a GPU: the channels-last UNet wiring, the hook layer, the step loops and
the frame-parallel partition, run in fp32 on CPU with the C-ABI kernel wrappers replaced by torch emulations
of their contracts (tests/cpu_ops_emulation.py, test-only).  Because nothing is bf16 here, the comparison with
the reference golden vectors and with the oracle is tight (1e-5; measured 3e-7) and isolates host-side wiring errors from
kernel numerics (which the `-m gpu` tests cover)."""
import copy
import os
import socket
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-5


def rel_l2(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.fixture()
def emulated_ops():
    from tests import cpu_ops_emulation as emu

    saved = emu.install()
    yield
    emu.uninstall(saved)


def _product_cpu(oracle_unet, kind):
    from mvoc_b200.unet3d import I2VGenXLUNet, UNetConfig, prepare

    m = I2VGenXLUNet(UNetConfig.named(kind)).eval().requires_grad_(False)
    m.load_state_dict(oracle_unet.state_dict(), strict=True)
    return prepare(m.to(torch.float32))


def _cond(inputs):
    from mvoc_b200.pipeline import Conditioning

    return Conditioning(inputs["prompt_embeds"], inputs["image_embeddings"], inputs["image_latents_first"],
                        inputs["image_latents"], inputs["fps"])


def _composite(wl, sched, inputs, unet, max_steps, parallel=None):
    from mvoc_b200.pipeline import I2VGenXLPipeline, LatentBank, init_pnp

    pipe = I2VGenXLPipeline(unet, "cpu", parallel=parallel)
    init_pnp(pipe, sched, wl)
    banks = [LatentBank(src, "cpu", pin_host=False) for src in inputs["source_latents"]]
    return pipe.sample_with_pnp_pipeline_with_edit_prompt_extraction_with_attn_injection(
        _cond(inputs), inputs["init_latents"].clone(), banks[0], banks[1:], list(inputs["masks"]),
        num_inference_steps=wl.n_steps, guidance_scale=wl.cfg, ddim_init_latents_t_idx=wl.ddim_init_latents_t_idx,
        fusion_steps=tuple(wl.fusion_step), random_noise_ratio=wl.random_noise_ratio,
        obj_random_noise_fusion=wl.obj_random_noise_fusion, max_steps=max_steps)


def _setup(name):
    from mvoc_b200 import synthetic
    from mvoc_b200.scheduler import DDIMSchedule
    from oracle import pipeline as opipe

    wl = synthetic.WORKLOADS[name]
    sched = DDIMSchedule(wl.n_steps)
    inputs = synthetic.make_inputs(wl, sched.timesteps, sched.alphas_cumprod)
    return wl, sched, inputs, opipe.build_unet(wl.unet, seed=0)


@pytest.mark.parametrize("case_name", ["all_hooks_t981", "attn_only_t481", "inject_bg_t481", "no_hooks_t21"])
def test_host_unet_forward_vs_reference_golden(emulated_ops, case_name):
    """Product host layer (fp32, emulated kernels) against the output of the REFERENCE'S OWN hook / UNet-driver
    code on the 4-level golden model (tests/golden/make_golden.py)."""
    from mvoc_b200 import pnp_utils
    from mvoc_b200.pipeline import I2VGenXLPipeline, init_pnp
    from mvoc_b200.scheduler import DDIMSchedule
    from tests.golden import spec

    gold = torch.load(os.path.join(ROOT, "tests", "golden", "unet_extension_forward_tiny4.pt"), map_location="cpu")
    case = next(c for c in spec.CASES if c["name"] == case_name)
    pu = _product_cpu(spec.build_tiny4(seed=0), "tiny4")
    pipe = I2VGenXLPipeline(pu, "cpu")
    cfg = SimpleNamespace(n_steps=50, pnp_f_t=case["pnp_f_t"], pnp_spatial_attn_t=case["pnp_spatial_attn_t"],
                          pnp_temp_attn_t=case["pnp_temp_attn_t"], inject_background=case["inject_background"])
    init_pnp(pipe, DDIMSchedule(50), cfg)
    inp = spec.make_inputs(case)
    pnp_utils.register_time_all(pipe, case["t"], list(inp["masks"]))
    with torch.no_grad():
        out = pipe._gather_prediction(pipe._unet_forward(inp["sample"], case["t"], _cond(inp)))
    err = rel_l2(out, gold[case_name])
    assert err <= TOL, f"{case_name}: {err:.3e}"


def test_host_composite_loop_vs_oracle(emulated_ops):
    """reduced2 (bg + 2 objects), 3 steps: fusion on step 0, feature + attention injection (the 10-step runs against
    the reference-generated vectors below cover the attention-only steps)."""
    from oracle import pipeline as opipe

    wl, sched, inputs, ou = _setup("reduced2")
    ref = opipe.composite_loop(copy.deepcopy(ou), wl, inputs, max_steps=3)
    out = _composite(wl, sched, inputs, _product_cpu(ou, wl.unet), 3)
    err = rel_l2(out, ref)
    assert err <= TOL, f"{err:.3e}"


def test_host_invert_loop_vs_oracle(emulated_ops, tmp_path):
    from mvoc_b200 import synthetic
    from mvoc_b200.pipeline import I2VGenXLPipeline, load_ddim_latents_at_t
    from oracle import pipeline as opipe

    wl = synthetic.WORKLOADS["config1"]
    inv = synthetic.make_inversion_inputs(wl)
    ou = opipe.build_unet(wl.unet, seed=0)
    ref = opipe.invert_loop(ou, wl, inv, max_steps=3)
    pipe = I2VGenXLPipeline(_product_cpu(ou, wl.unet), "cpu")
    saved = pipe.invert(inv["latents"].clone(), inv["prompt_embeds"], inv["image_embeddings"], inv["image_latents"],
                        inv["fps"], num_inference_steps=wl.inversion_steps, output_dir=str(tmp_path), max_steps=3)
    assert sorted(saved) == sorted(ref) == [1, 3, 5]
    for t in saved:
        assert rel_l2(saved[t], ref[t]) <= TOL
        # the file holds the reference's wire dtype (its fp16 pipeline latents, pipeline_i2vgen_xl.py:1990-1993)
        on_disk = load_ddim_latents_at_t(t, str(tmp_path))
        assert on_disk.dtype == torch.float16 and torch.equal(on_disk, saved[t].to(torch.float16))


# ------------------------------------------------------------------ frame-parallel (gloo, world_size 2)
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fp_setup(variant):
    """(workload, schedule, inputs, oracle UNet) of a frame-parallel scenario."""
    if variant == "reduced2":
        return _setup("reduced2")
    # 4-level model on 24x24 latents: the lowest level has 3x3 = 9 pixels, which no even world size divides ->
    # its temporal operators run replicated on all-gathered frames (FrameParallel._temporal_replicated)
    from mvoc_b200 import synthetic
    from mvoc_b200.scheduler import DDIMSchedule
    from tests.golden import spec

    wl = (synthetic.Workload("tiny4_24", "tiny4", 4, 24, 24, 2) if variant == "tiny4_24"
          else synthetic.Workload("tiny4_t8", "tiny4", 8, 32, 32, 2))       # 8 ranks: one frame each, h == world at 8x8
    sched = DDIMSchedule(wl.n_steps)
    return wl, sched, synthetic.make_inputs(wl, sched.timesteps, sched.alphas_cumprod), spec.build_tiny4(seed=0)


def _fp_worker(rank, world, port, ret, variant="reduced2", gather_max_pixels=0):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mvoc_b200.parallel import FrameParallel
        from tests import cpu_ops_emulation as emu

        emu.install()
        FrameParallel._GATHER_MAX_PIXELS = gather_max_pixels
        wl, sched, inputs, ou = _fp_setup(variant)
        par = FrameParallel(dist.group.WORLD, world, rank, torch.device("cpu"))
        with torch.no_grad():
            sharded = _composite(wl, sched, inputs, _product_cpu(ou, wl.unet), 2, parallel=par)
            single = _composite(wl, sched, inputs, _product_cpu(ou, wl.unet), 2) if rank == 0 else torch.empty_like(sharded)
        dist.broadcast(single, src=0)           # the single-rank reference is computed once, by rank 0
        err = rel_l2(sharded, single)
        # every rank holds the full updated latents (the prediction is all-gathered before the DDIM update)
        got = [torch.empty_like(sharded) for _ in range(world)]
        dist.all_gather(got, sharded)
        same = all(torch.equal(g, got[0]) for g in got)
        ret[rank] = ("ok", err, same)
    except Exception:
        import traceback

        ret[rank] = ("error", traceback.format_exc(), False)
    finally:
        dist.destroy_process_group()


def _run_fp(world, variant="reduced2", gather_max_pixels=0):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fp_worker, args=(world, _free_port(), ret, variant, gather_max_pixels), nprocs=world, join=True)
    for r in range(world):
        status, err, same = ret.get(r)
        assert status == "ok", f"rank {r}: {err}"
        assert err <= TOL, f"rank {r}: sharded vs single {err:.3e}"
        assert same, "ranks disagree on the updated latents"


def test_host_frame_parallel_indivisible_level_is_replicated():
    """A level whose pixel count no rank count divides (3x3 here; config 5's 11x20 on 8 GPUs) runs its temporal
    operators on all-gathered frames instead of pixel shards — same latents as a single rank."""
    _run_fp(2, "tiny4_24")


@pytest.mark.skipif(os.environ.get("MVOC_LONG_TESTS") != "1", reason="optional switch; ~25 s")
def test_host_frame_parallel_gather_threshold():
    """MVOC_FP_GATHER_MAX_PIXELS: low-resolution levels (<= 256 pixels here: the 16x16 level of reduced2)
    replicated by choice, the 32x32 level pixel-sharded."""
    _run_fp(2, "reduced2", gather_max_pixels=256)


@pytest.mark.parametrize("world", [2, pytest.param(4, marks=pytest.mark.skipif(
    os.environ.get("MVOC_LONG_TESTS") != "1", reason="4 processes; ~30 s (world 2 runs by default, 8 is long too)"))])
def test_host_composite_frame_parallel(world):
    """P ranks, each running every branch on 1/P of the frames (temporal operators on pixel shards after the
    all-to-all, GroupNorm statistics merged across shards): same latents as the single-rank loop."""
    _run_fp(world)


def test_host_dense_routing_vs_oracle(emulated_ops):
    """The tcgen05 GEMM family is the product path: tap-major 3x3 / temporal filters, fused QKV, bias + residual,
    the 1x1 shortcut conv inside conv2 and the GEGLU gate all ride in one call each.  With the kernels emulated in
    torch the composition result equals the oracle, and every fusion is exercised."""
    from oracle import pipeline as opipe
    from tests import cpu_ops_emulation as emu

    wl, sched, inputs, ou = _setup("reduced2")
    ref = opipe.composite_loop(copy.deepcopy(ou), wl, inputs, max_steps=1)
    out = _composite(wl, sched, inputs, _product_cpu(ou, wl.unet), 1)
    assert rel_l2(out, ref) <= TOL
    for key in ("linear", "linear_res", "geglu", "conv", "conv_res", "conv_shortcut", "tconv", "tconv_res"):
        assert emu.calls[key] > 0, key


def test_host_library_dense_path_vs_oracle():
    """MVOC_DENSE=lib (the cuDNN / cuBLAS A/B baseline of bench.py) gives the same composition result."""
    from oracle import pipeline as opipe
    from tests import cpu_ops_emulation as emu

    saved = emu.install(dense=False)
    try:
        wl, sched, inputs, ou = _setup("reduced2")
        ref = opipe.composite_loop(copy.deepcopy(ou), wl, inputs, max_steps=1)
        out = _composite(wl, sched, inputs, _product_cpu(ou, wl.unet), 1)
        assert rel_l2(out, ref) <= TOL
        assert emu.calls["linear"] == 0 and emu.calls["conv"] == 0
    finally:
        emu.uninstall(saved)


def test_derived_weights_follow_load_state_dict(emulated_ops):
    """Cached weight re-layouts (tap-major filters, fused QKV) are rebuilt after load_state_dict (advisor r1)."""
    from mvoc_b200.unet3d import I2VGenXLUNet, UNetConfig, derived

    m = I2VGenXLUNet(UNetConfig.named("reduced")).eval().requires_grad_(False)
    attn = m.down_blocks[0].attentions[0].transformer_blocks[0].attn1
    x = torch.randn(2, 4, attn.to_q.in_features)
    q0 = attn.qkv_self(x)[0].clone()
    epoch0 = m.__dict__.get("derived_epoch", 0)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    for k in sd:
        if k.endswith("attn1.to_q.weight"):
            sd[k] = sd[k] * 2.0
    m.load_state_dict(sd)
    assert m.__dict__["derived_epoch"] == epoch0 + 1
    assert torch.allclose(attn.qkv_self(x)[0], 2.0 * q0, atol=1e-5)
    attn.to_q.weight.mul_(0.5)                    # an in-place edit without load_state_dict is seen as well
    assert torch.allclose(attn.qkv_self(x)[0], q0, atol=1e-5)


# ------------------------------------------------------------------ the reference's own step loops (golden)
@pytest.mark.parametrize("case", ["default", "exotic"])
def test_host_composition_loop_vs_reference_golden(emulated_ops, case):
    """Product loop (host layer in fp32, kernels emulated) against the latents the REFERENCE'S OWN composition loop
    produced (tests/golden/make_golden_loops.py) — every loop option off its default in the `exotic` case."""
    from mvoc_b200.pipeline import Conditioning, I2VGenXLPipeline, LatentBank, init_pnp
    from mvoc_b200.scheduler import DDIMSchedule
    from tests.golden import spec

    gold = torch.load(os.path.join(ROOT, "tests", "golden", "composition_loop_tiny4.pt"), map_location="cpu")
    fx = spec.loop_fixture(case)
    wl = spec.loop_workload(fx)
    inp = spec.loop_inputs(fx, gold["seam"])
    sched = DDIMSchedule(fx["n_steps"])
    pipe = I2VGenXLPipeline(_product_cpu(spec.build_tiny4(seed=0), "tiny4"), "cpu")
    init_pnp(pipe, sched, wl)
    banks = [LatentBank(src, "cpu", pin_host=False) for src in inp["source_latents"]]
    rec = {}
    pipe.sample_with_pnp_pipeline_with_edit_prompt_extraction_with_attn_injection(
        Conditioning(inp["prompt_embeds"], inp["image_embeddings"], inp["image_latents_first"], inp["image_latents"],
                     inp["fps"]),
        inp["init_latents"].clone(), banks[0], banks[1:], inp["masks"], num_inference_steps=fx["n_steps"],
        guidance_scale=fx["cfg"], ddim_init_latents_t_idx=fx["ddim_init_latents_t_idx"],
        fusion_steps=tuple(fx["fusion_step"]), random_noise_ratio=fx["random_noise_ratio"],
        obj_random_noise_fusion=fx["obj_random_noise_fusion"],
        obj_ddim_latents_idx_offset=fx["obj_ddim_latents_idx_offset"], max_steps=10,
        callback=lambda i, t, lat: rec.__setitem__(i, lat.clone()))
    checked = 0
    for i, ref in gold["latents_after_step"][case].items():
        if i in rec:
            assert rel_l2(rec[i], ref) <= TOL, f"{case} step {i}: {rel_l2(rec[i], ref):.3e}"
            checked += 1
    assert checked >= 5
    # the same as PSNR of decoded frames, one oracle-side decoder for both latents
    from oracle import vae

    dec = vae.build_decoder()
    db = vae.psnr(vae.decode_latents(dec, rec[9]), vae.decode_latents(dec, gold["latents_after_step"][case][9]))
    assert db >= 100.0, db


def test_host_inversion_loop_vs_reference_golden(emulated_ops):
    from mvoc_b200.pipeline import I2VGenXLPipeline
    from tests.golden import spec

    gold = torch.load(os.path.join(ROOT, "tests", "golden", "inversion_loop_tiny4.pt"), map_location="cpu")
    ix = spec.inversion_fixture()
    seam = gold["seam"]
    pipe = I2VGenXLPipeline(_product_cpu(spec.build_tiny4(seed=0), "tiny4"), "cpu")
    saved = pipe.invert(spec.inversion_init_latents(ix), seam["encoder_hidden_states"], seam["image_embeddings"],
                        seam["image_latents"], seam["fps"], num_inference_steps=ix["n_steps"], max_steps=3)
    assert sorted(saved) == [1, 3, 5]
    for t in saved:
        assert rel_l2(saved[t], gold["latents_at_t"][t]) <= TOL


@pytest.mark.skipif(os.environ.get("MVOC_LONG_TESTS") != "1", reason="1.4 B-parameter model twice in fp32: ~12 GB, minutes")
def test_host_full_architecture_forward_vs_oracle(emulated_ops):
    """The FULL i2vgen-xl architecture (320/640/1280/1280, 5/10/20 heads — the benchmark's model) through the
    product host layer against the oracle, one composite step on 2 frames x 16x16 latents, every hook firing."""
    from mvoc_b200 import synthetic
    from mvoc_b200.scheduler import DDIMSchedule
    from oracle import pipeline as opipe

    wl = synthetic.Workload("full_tiny", "full", 2, 16, 16, 2)
    sched = DDIMSchedule(wl.n_steps)
    inputs = synthetic.make_inputs(wl, sched.timesteps, sched.alphas_cumprod)
    ou = opipe.build_unet("full", seed=0)
    with torch.no_grad():
        ref = opipe.composite_loop(ou, wl, inputs, max_steps=1)
        pu = _product_cpu(ou, "full")
        del ou
        out = _composite(wl, sched, inputs, pu, 1)
    assert rel_l2(out, ref) <= TOL


@pytest.mark.skipif(os.environ.get("MVOC_LONG_TESTS") != "1", reason="8 processes; about a minute")
def test_host_composite_frame_parallel_world8():
    """Eight ranks with one frame each on the 4-level model (the 8x8 level has h == world_size, the layout that
    once produced a non-contiguous relayout on the 8-GPU box)."""
    _run_fp(8, "tiny4_t8")
