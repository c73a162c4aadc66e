"""Parity of the BENCHMARKED configuration and of the full-length step loops on the GPU.

* the full 1.42 B-parameter UNet at 64x64 latents, background + 2 objects (BASELINE config 2 with 2 of its 16
  frames, so that the fp32 CPU oracle can pay for it): one UNet forward with every hook firing inside the
  1.5x torch-bf16 envelope, and one whole composition step (fusion, feature + attention injection, CFG, DDIM);
* all 50 steps of the reference's own composition loop and all 500 steps of its inversion loop (golden vectors
  produced by executing /root/reference code, tests/golden/make_golden_loops.py) with STATED final bars:
  relative L2 of the final latents and PSNR of the frames decoded from them by one oracle-side decoder.
"""
import copy
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Bars (bf16 kernels vs fp32 reference), each ~2x what B200 measured (profiles/r02_pytest_gpu_b.log):
# one forward: 1.5x the torch-bf16 envelope measured in the same test (1.40e-2 vs envelope 1.59e-2);
# one composition step: 1.6e-4 (the DDIM update damps the prediction error);
# 50 steps: final latents 1.15e-2 / 1.02e-2, decoded-frame PSNR 56.5 / 58.9 dB; 500 inversion steps: 4.7e-3 at t=999.
BAR_STEP = 5e-3
BAR_LOOP50 = 2.5e-2
PSNR_LOOP50 = 50.0
BAR_INV500 = 1.5e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


def test_full_unet_config2_vs_oracle(cuda_device):
    """UNetConfig.full() at 64x64, 5 branches: forward (all hooks) inside the bf16 envelope + one composite step."""
    from mvoc_b200 import pnp_utils, synthetic
    from mvoc_b200.pipeline import I2VGenXLPipeline, init_pnp
    from mvoc_b200.scheduler import DDIMSchedule
    from oracle import hooks as ohooks
    from oracle import pipeline as opipe
    from tests.test_pipeline_gpu import _cond, _product_from, _run_product_composite

    wl = synthetic.WORKLOADS["config2_t2"]
    sched = DDIMSchedule(wl.n_steps)
    inputs = synthetic.make_inputs(wl, sched.timesteps, sched.alphas_cumprod)
    ou = opipe.build_unet(wl.unet, seed=0)
    assert sum(p.numel() for p in ou.parameters()) == 1_420_469_224
    t = sched.timesteps[0]
    torch.manual_seed(11)
    sample = torch.randn(wl.n_branches, 4, wl.n_frames, wl.latent_h, wl.latent_w)
    # whole step on the fp32 oracle first (it owns `ou` un-hooked through a deepcopy)
    ref_step = opipe.composite_loop(copy.deepcopy(ou), wl, inputs, max_steps=1)
    ou16 = copy.deepcopy(ou).to(cuda_device, torch.bfloat16)
    pu = _product_from(ou, wl, cuda_device)
    ns = SimpleNamespace(unet=ou)
    opipe.init_pnp(ns, torch.tensor(sched.timesteps), wl)
    ohooks.register_time_all(ns, t, inputs["masks"])
    ref = opipe.unet_extension_forward(ou, sample, t, inputs["fps"], inputs["image_latents_first"],
                                       inputs["image_latents"], inputs["image_embeddings"], inputs["prompt_embeds"])
    ns16 = SimpleNamespace(unet=ou16)
    opipe.init_pnp(ns16, torch.tensor(sched.timesteps), wl)
    masks_d = [(mf.to(cuda_device), mb.to(cuda_device)) for mf, mb in inputs["masks"]]
    ohooks.register_time_all(ns16, t, masks_d)
    bf = lambda x: x.to(cuda_device, torch.bfloat16)
    env = opipe.unet_extension_forward(ou16, bf(sample), t, inputs["fps"].to(cuda_device),
                                       bf(inputs["image_latents_first"]), bf(inputs["image_latents"]),
                                       bf(inputs["image_embeddings"]), bf(inputs["prompt_embeds"]))
    env_err = rel_l2(env, ref)
    del ou16, ns16, env
    torch.cuda.empty_cache()
    pipe = I2VGenXLPipeline(pu, cuda_device)
    init_pnp(pipe, sched, wl)
    pnp_utils.register_time_all(pipe, t, masks_d)
    assert all(pnp_utils.hook_signature(pu)), "every hook must fire at the first timestep"
    out = pipe._gather_prediction(pipe._unet_forward(bf(sample), t, _cond(inputs, cuda_device)))
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    comp = rel_l2(out[wl.n_obj + 1:], ref[wl.n_obj + 1:])
    print(f"[config2 model, T=2] forward rel L2 {err:.4e} (composite branches {comp:.4e}); "
          f"torch-bf16 envelope {env_err:.4e}")
    assert torch.isfinite(out.float()).all()
    assert err <= 1.5 * env_err + 2e-3, f"product {err:.3e} vs envelope {env_err:.3e}"
    assert comp <= 1.5 * env_err + 4e-3
    step = _run_product_composite(wl, sched, inputs, pu, cuda_device, 1)
    e_step = rel_l2(step, ref_step)
    print(f"[config2 model, T=2] one composite step: latents rel L2 {e_step:.4e} (bar {BAR_STEP:.0e})")
    assert e_step <= BAR_STEP


def _tiny4(device):
    from mvoc_b200.unet3d import I2VGenXLUNet, UNetConfig, prepare
    from tests.golden import spec

    m = I2VGenXLUNet(UNetConfig.tiny4()).eval().requires_grad_(False)
    m.load_state_dict(spec.build_tiny4(seed=0).state_dict(), strict=True)
    return prepare(m.to(device, torch.bfloat16))


@pytest.mark.parametrize("case", ["default", "exotic"])
def test_composition_all_50_steps_vs_reference_golden(cuda_device, case):
    """The reference's own 50-step loop (pipeline_i2vgen_xl.py:1636-1734) end to end, CUDA graphs on."""
    from mvoc_b200.pipeline import Conditioning, I2VGenXLPipeline, LatentBank, init_pnp
    from mvoc_b200.scheduler import DDIMSchedule
    from oracle import vae
    from tests.golden import spec

    gold = torch.load(os.path.join(GOLDEN, "composition_loop_tiny4.pt"), map_location="cpu")
    fx = spec.loop_fixture(case)
    inp = spec.loop_inputs(fx, gold["seam"])
    pipe = I2VGenXLPipeline(_tiny4(cuda_device), cuda_device, use_cuda_graphs=True)
    init_pnp(pipe, DDIMSchedule(fx["n_steps"]), spec.loop_workload(fx))
    bf = lambda x: x.to(cuda_device, torch.bfloat16)
    banks = [LatentBank(src, cuda_device, pin_host=False) for src in inp["source_latents"]]
    masks = [(mf.to(cuda_device), mb.to(cuda_device)) for mf, mb in inp["masks"]]
    rec = {}
    pipe.sample_with_pnp_pipeline_with_edit_prompt_extraction_with_attn_injection(
        Conditioning(bf(inp["prompt_embeds"]), bf(inp["image_embeddings"]), bf(inp["image_latents_first"]),
                     bf(inp["image_latents"]), inp["fps"].to(cuda_device)),
        inp["init_latents"].to(cuda_device).clone(), banks[0], banks[1:], masks, num_inference_steps=fx["n_steps"],
        guidance_scale=fx["cfg"], ddim_init_latents_t_idx=fx["ddim_init_latents_t_idx"],
        fusion_steps=tuple(fx["fusion_step"]), random_noise_ratio=fx["random_noise_ratio"],
        obj_random_noise_fusion=fx["obj_random_noise_fusion"],
        obj_ddim_latents_idx_offset=fx["obj_ddim_latents_idx_offset"],
        callback=lambda i, t, lat: rec.__setitem__(i, lat.detach().float().cpu()))
    torch.cuda.synchronize()
    want = gold["latents_after_step"][case]
    last = max(want)
    assert last in rec and last >= 48, (last, sorted(rec)[-3:])
    errs = {i: rel_l2(rec[i], ref) for i, ref in want.items()}
    dec = vae.build_decoder()
    db = vae.psnr(vae.decode_latents(dec, rec[last]), vae.decode_latents(dec, want[last]))
    print(f"[reference loop, {case}, all steps] rel L2 per golden step {errs}; PSNR after step {last}: {db:.1f} dB")
    assert errs[last] <= BAR_LOOP50, errs
    assert db >= PSNR_LOOP50, db


def test_inversion_all_500_steps_vs_reference_golden(cuda_device):
    """The reference's own 500-step DDIM inversion (pipeline_i2vgen_xl.py:1940-2000) end to end, CUDA graphs on."""
    from mvoc_b200.pipeline import I2VGenXLPipeline
    from tests.golden import spec

    gold = torch.load(os.path.join(GOLDEN, "inversion_loop_tiny4.pt"), map_location="cpu")
    ix = spec.inversion_fixture()
    seam = gold["seam"]
    bf = lambda x: x.to(cuda_device, torch.bfloat16)
    pipe = I2VGenXLPipeline(_tiny4(cuda_device), cuda_device, use_cuda_graphs=True)
    saved = pipe.invert(spec.inversion_init_latents(ix).to(cuda_device), bf(seam["encoder_hidden_states"]),
                        bf(seam["image_embeddings"]), bf(seam["image_latents"]), seam["fps"].to(cuda_device),
                        num_inference_steps=ix["n_steps"])
    torch.cuda.synchronize()
    assert len(saved) == ix["n_steps"]
    errs = {t: rel_l2(saved[t], ref) for t, ref in gold["latents_at_t"].items()}
    print(f"[reference inversion, all steps] rel L2 at golden timesteps {errs}")
    assert max(errs.values()) <= BAR_INV500, errs
