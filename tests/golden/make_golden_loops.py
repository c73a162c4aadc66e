"""Generates tests/golden/{composition,inversion}_loop_tiny4.pt by running the REFERENCE'S OWN step loops
(run in the build container only; needs /root/reference):

    python tests/golden/make_golden_loops.py

Executed from /root/reference, unmodified:
  * pipelines/pipeline_i2vgen_xl.py — I2VGenXLPipeline.sample_with_pnp_pipeline_with_edit_prompt_extraction_
    with_attn_injection (:1220-1750: conditioning assembly, timestep / fusion-timestep selection, per-step latent
    loads, noise fusion, branch concat, CFG, reshape + scheduler step), I2VGenXLPipeline.invert (:1752-2003),
    prepare_image_latents (:860-890), prepare_latents, prepare_extra_step_kwargs, do_classifier_free_guidance,
    I2VGenXLUnetExtension.forward, _center_crop_wide / _resize_bilinear
  * composite.py — init_pnp;  pnp_utils.py — every register_* hook;  utils.py — mask_preprocess (on the PNGs of
    tests/golden/masks128), load_ddim_latents_at_t (on files this script writes to a temp directory)
The pipeline object is created without __init__ and given test doubles for everything outside the hot path:
CLIP text / vision encoders (`encode_prompt`, `_encode_image`: seeded tensors), the VAE (`vae.encode`: average
pooling), `image_processor.preprocess`, the progress bar; the UNet is the oracle's 4-level "tiny4" model and the
scheduler is the oracle's DDIM restatement behind the diffusers call interface (diffusers itself is un-vendored,
see oracle/unet.py, oracle/scheduler.py).  What these vectors pin is therefore the MVOC step loops themselves.

Stored: the tensors at the seam (what the loop hands to the UNet: prompt / image embeddings, image latents, fps),
and the latents after selected steps.  Source latents, initial noise and images are regenerated from seeds
(tests/golden/spec.py); the mask PNGs are committed.
"""
from __future__ import annotations

import os
import sys
import tempfile
from copy import deepcopy
from functools import partial
from types import SimpleNamespace

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import scheduler as osched  # noqa: E402
from tests.golden import spec  # noqa: E402
from tests.golden.make_golden import REF, install_stubs, load_ref  # noqa: E402


# ------------------------------------------------------------------ test doubles (outside the hot path)
class _Dist:
    def __init__(self, x):
        self.x = x

    def sample(self, generator=None):
        return self.x


class FakeVAE:
    config = SimpleNamespace(scaling_factor=0.18215)

    def encode(self, image):
        p = F.avg_pool2d(image.float(), 8)
        return SimpleNamespace(latent_dist=_Dist(torch.cat([p, p.mean(1, keepdim=True)], dim=1) * 3.0))


class FakeImageProcessor:
    def preprocess(self, pil):
        import numpy as np

        a = torch.from_numpy(np.asarray(pil.convert("RGB"), dtype=np.float32).copy())
        return (a / 127.5 - 1.0).permute(2, 0, 1)[None]


class SchedulerAdapter:
    """oracle DDIM restatement behind diffusers' scheduler interface; records every prev_sample."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, inverse: bool = False):
        self.core = osched.DDIMInverseScheduler() if inverse else osched.DDIMScheduler()
        self.timesteps = None
        self.record = []

    def set_timesteps(self, n, device=None):
        self.core.set_timesteps(n)
        self.timesteps = self.core.timesteps.clone()

    def scale_model_input(self, x, t):
        return x

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None):
        assert eta == 0.0
        prev = self.core.step(model_output, int(timestep), sample)
        self.record.append(prev.clone())
        return SimpleNamespace(prev_sample=prev)


def _to5d(x, fx):
    return x[None, :].reshape(1, fx["T"], 4, fx["H"], fx["W"]).permute(0, 2, 1, 3, 4).contiguous()


class _Bar:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def update(self):
        pass


def _pipeline_config(unet):
    """The two config fields the pipeline reads besides the oracle's own (sample_size is only a default)."""
    for k, v in (("sample_size", 16), ("in_channels", 4)):
        if not hasattr(unet.config, k):
            object.__setattr__(unet.config, k, v)


def make_pipe(ref_pipe, unet, scheduler, guidance_scale):
    pipe = object.__new__(ref_pipe.I2VGenXLPipeline)          # reference class, no diffusers __init__
    pipe.unet = unet
    pipe.vae = FakeVAE()
    pipe.vae_scale_factor = 8
    pipe.image_processor = FakeImageProcessor()
    pipe.feature_extractor = SimpleNamespace(crop_size={"width": 32, "height": 32})
    pipe.scheduler = scheduler
    pipe._guidance_scale = guidance_scale
    pipe.check_inputs = lambda *a, **k: None
    pipe.progress_bar = lambda total=None: _Bar()
    cls = type(pipe)
    cls._execution_device = property(lambda self: torch.device("cpu"))
    cls.device = property(lambda self: torch.device("cpu"))

    def encode_prompt(prompt, device, n, negative_prompt=None, prompt_embeds=None, negative_prompt_embeds=None, **k):
        return (spec.text_embedding(prompt) if prompt_embeds is None else prompt_embeds,
                spec.text_embedding(negative_prompt or "") if negative_prompt_embeds is None else negative_prompt_embeds)

    def _encode_image(image, device, n):
        emb = spec.clip_embedding(image)[None, None]                       # [1, 1, 1024]
        if pipe.do_classifier_free_guidance:                               # as the real method (:764-768)
            emb = torch.cat([torch.zeros_like(emb), emb])
        return emb

    pipe.encode_prompt = encode_prompt
    pipe._encode_image = _encode_image
    return pipe


def run_composition(ref_pipe, ref_comp, fx):
    unet = spec.build_tiny4(seed=0)
    _pipeline_config(unet)
    sched = SchedulerAdapter()
    sched.set_timesteps(fx["n_steps"])
    pipe = make_pipe(ref_pipe, unet, sched, fx["cfg"])
    cfg = SimpleNamespace(n_steps=fx["n_steps"], pnp_f_t=fx["pnp_f_t"], pnp_spatial_attn_t=fx["pnp_spatial_attn_t"],
                          pnp_temp_attn_t=fx["pnp_temp_attn_t"], pnp_cross_attn_t=0.0,
                          inject_background=fx["inject_background"])
    ref_comp.init_pnp(pipe, sched, cfg)                                                   # composite.py:157
    unet.forward = partial(ref_pipe.I2VGenXLUnetExtension.forward, unet)                  # composite.py:163
    seam = {}
    inner = unet.forward

    def spy(sample, t, **kw):
        if not seam:
            seam.update({k: v.clone() for k, v in kw.items() if isinstance(v, torch.Tensor)})
        return inner(sample, t, **kw)

    unet.forward = spy
    with tempfile.TemporaryDirectory() as tmp:
        dirs = spec.write_source_latents(tmp, fx)                                          # ddim_latents_{t}.pt
        imgs = spec.loop_images(fx)
        out = pipe.sample_with_pnp_pipeline_with_edit_prompt_extraction_with_attn_injection(
            prompt=fx["prompt"], main_first_image=imgs["main"][0], main_image_list=imgs["main"],
            background_first_image=imgs["bg"][0], background_image_list=imgs["bg"],
            objs_first_image=[o[0] for o in imgs["objs"]], objs_image_list=imgs["objs"],
            height=fx["H"] * 8, width=fx["W"] * 8, num_frames=fx["T"], num_inference_steps=fx["n_steps"],
            guidance_scale=fx["cfg"], negative_prompt=fx["negative_prompt"], target_fps=fx["fps"],
            latents=spec.loop_init_latents(fx), output_type="latent", return_dict=True,
            ddim_init_latents_t_idx=fx["ddim_init_latents_t_idx"], ddim_inv_prompt=fx["ddim_inv_prompt"],
            obj_mask=spec.mask_paths(os.path.join(HERE, "masks128")), obj_width_height=[(0, 0)] * fx["n_obj"],
            random_noise_ratio=fx["random_noise_ratio"], bg_inv_latents_path=dirs[0],
            obj_ddim_latents_path=dirs[1:], obj_ddim_latents_idx_offset=fx["obj_ddim_latents_idx_offset"],
            obj_random_noise_fusion=fx["obj_random_noise_fusion"], fusion_steps=fx["fusion_step"]).frames
    to5d = lambda x: _to5d(x, fx)
    steps = {i: to5d(sched.record[i]) for i in fx["keep_steps"]}
    assert torch.equal(steps[fx["keep_steps"][-1]], out) and len(sched.record) == fx["n_steps"] - fx["ddim_init_latents_t_idx"]
    print("composition", fx["case"], {k: tuple(v.shape) for k, v in seam.items()}, "final std", float(out.std()))
    return {"seam": seam, "steps": steps}



def main():
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.dirname(REF))
    load_ref("utils", "utils.py")
    load_ref("pnp_utils", "pnp_utils.py")
    ref_pipe = load_ref("ref_pipeline_i2vgen_xl", "pipelines/pipeline_i2vgen_xl.py")
    ref_comp = load_ref("ref_composite", "composite.py")
    torch.set_grad_enabled(False)
    spec.write_mask_pngs(os.path.join(HERE, "masks128"))

    # ---------------------------------------------------------------- composition (composite.py:157-200)
    comp = {}
    for case in spec.LOOP_CASES:
        comp[case] = run_composition(ref_pipe, ref_comp, spec.loop_fixture(case))
    seam = comp["default"]["seam"]
    for c in comp.values():   # the conditioning does not depend on the loop options (fps / cfg aside)
        assert all(torch.equal(c["seam"][k], seam[k]) for k in seam)
    torch.save({"seam": seam, "latents_after_step": {c: v["steps"] for c, v in comp.items()}},
               os.path.join(HERE, "composition_loop_tiny4.pt"))

    # ---------------------------------------------------------------- inversion (inverse.py:48-76)
    ix = spec.inversion_fixture()
    unet = spec.build_tiny4(seed=0)                       # stock forward: diffusers' (= the oracle's restatement)
    _pipeline_config(unet)
    sched = SchedulerAdapter(inverse=True)
    pipe = make_pipe(ref_pipe, unet, sched, ix["cfg"])
    seam = {}
    inner = unet.forward

    def spy2(sample, t, **kw):
        if not seam:
            seam.update({k: v.clone() for k, v in kw.items() if isinstance(v, torch.Tensor)})
        return inner(sample, t, **kw)

    unet.forward = spy2
    with tempfile.TemporaryDirectory() as tmp:
        stacked = pipe.invert(prompt=ix["prompt"], image=spec.loop_images(ix)["main"][0], height=ix["H"] * 8,
                              width=ix["W"] * 8, num_frames=ix["T"], num_inference_steps=ix["n_steps"],
                              guidance_scale=ix["cfg"], negative_prompt=ix["negative_prompt"], target_fps=ix["fps"],
                              latents=spec.inversion_init_latents(ix), output_dir=tmp, return_dict=False)
        files = sorted(os.listdir(tmp))
        ts = [int(t) for t in sched.timesteps]
        assert len(files) == ix["n_steps"] and f"ddim_latents_{ts[0]}.pt" in files
        disk = {t: torch.load(os.path.join(tmp, f"ddim_latents_{t}.pt")) for t in ix["keep_timesteps"]}
    to5d = lambda x: _to5d(x, ix)
    for t in ix["keep_timesteps"]:
        assert torch.equal(disk[t], to5d(sched.record[ts.index(t)]))
    assert stacked.shape[1] == ix["n_steps"]
    torch.save({"seam": seam, "latents_at_t": disk}, os.path.join(HERE, "inversion_loop_tiny4.pt"))
    print("inversion:", {k: tuple(v.shape) for k, v in seam.items()}, "kept", sorted(disk))


if __name__ == "__main__":
    main()
