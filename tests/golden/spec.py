"""Shared definition of the golden cases (inputs are regenerated from seeds; only outputs are stored)."""
from __future__ import annotations

import torch

from mvoc_b200 import synthetic
from oracle.unet import I2VGenXLUNet, UNetConfig

T, H, W, N_OBJ = 4, 16, 16, 2

CASES = [
    dict(name="all_hooks_t981", t=981, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0, inject_background=False),
    dict(name="attn_only_t481", t=481, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0, inject_background=False),
    dict(name="inject_bg_t481", t=481, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0, inject_background=True),
    dict(name="no_hooks_t21", t=21, pnp_f_t=0.2, pnp_spatial_attn_t=0.2, pnp_temp_attn_t=0.5, inject_background=False),
]


def tiny4_config() -> UNetConfig:
    """Full 4-level I2VGenXL layout (what the reference's hard-coded indices need), narrow channels."""
    return UNetConfig(block_out_channels=(64, 128, 256, 256), transformer_in_heads=2)


def build_tiny4(seed: int = 0) -> I2VGenXLUNet:
    torch.manual_seed(seed)
    return I2VGenXLUNet(tiny4_config()).eval().requires_grad_(False)


def timesteps_50():
    return [981 - 20 * i for i in range(50)]


def make_inputs(case) -> dict:
    wl = synthetic.Workload("tiny4", "tiny4", T, H, W, N_OBJ)
    base = synthetic.make_inputs(wl, [case["t"]], {case["t"]: 0.5})
    g = torch.Generator().manual_seed(77)
    base["sample"] = torch.randn(N_OBJ + 3, 4, T, H, W, generator=g)
    return base


# ---------------------------------------------------------------------------------------------------------
# Step-loop fixtures (make_golden_loops.py runs the reference's own composition / inversion loops on these)
# ---------------------------------------------------------------------------------------------------------
def _loop_common() -> dict:
    return dict(T=T, H=H, W=W, n_obj=N_OBJ, fps=8, prompt="a boat and a surfer on the sea", negative_prompt="blurry",
                ddim_inv_prompt="")


LOOP_CASES = {
    # boat_surf entry of the reference's group_config.json: fusion on the first step only, no noise mixing
    "default": dict(n_steps=50, cfg=9.0, pnp_f_t=0.1, pnp_spatial_attn_t=1.0, pnp_temp_attn_t=1.0,
                    inject_background=False, fusion_step=(0, 1), random_noise_ratio=0.0, obj_random_noise_fusion=False,
                    obj_ddim_latents_idx_offset=[0, 0], ddim_init_latents_t_idx=0, keep_steps=[0, 1, 2, 5, 9, 49]),
    # every option off its default: late start, two fusion steps (fusion_counter is never incremented, :1634),
    # per-object timestep offsets, noise mixing inside the object region, background injection
    "exotic": dict(n_steps=50, cfg=7.5, pnp_f_t=0.2, pnp_spatial_attn_t=0.6, pnp_temp_attn_t=0.4,
                   inject_background=True, fusion_step=(0, 2), random_noise_ratio=0.3, obj_random_noise_fusion=True,
                   obj_ddim_latents_idx_offset=[0, 2], ddim_init_latents_t_idx=1, keep_steps=[0, 1, 2, 5, 9, 48]),
}


def loop_fixture(case: str = "default") -> dict:
    return {**_loop_common(), **LOOP_CASES[case], "case": case}


def inversion_fixture() -> dict:
    return {**_loop_common(), "n_steps": 500, "cfg": 1.0, "prompt": "", "negative_prompt": None,
            "keep_timesteps": [1, 3, 5, 499, 999]}


def text_embedding(text: str) -> torch.Tensor:
    """Stand-in for the CLIP text encoder: [1, 8, 1024] seeded by the string."""
    seed = sum((i + 1) * ord(c) for i, c in enumerate(text)) % (2 ** 31 - 1)
    return torch.randn(1, 8, 1024, generator=torch.Generator().manual_seed(1234 + seed))


def clip_embedding(pil) -> torch.Tensor:
    """Stand-in for the CLIP vision encoder on a (resized) PIL frame: [1024]."""
    import numpy as np

    a = torch.from_numpy(np.asarray(pil.convert("RGB"), dtype=np.float32).copy()).permute(2, 0, 1) / 255.0
    feat = torch.nn.functional.adaptive_avg_pool2d(a[None], 8).reshape(-1)                     # 192
    proj = torch.randn(192, 1024, generator=torch.Generator().manual_seed(4242)) * 0.2
    return (feat - 0.5) @ proj


def loop_images(fx) -> dict:
    """Deterministic RGB frames (PIL, 8*W x 8*H) for the main / background / object branches."""
    import numpy as np
    from PIL import Image

    def frames(seed):
        out = []
        for i in range(fx["T"]):
            rng = np.random.RandomState(seed * 100 + i)
            small = rng.randint(0, 256, size=(8, 8, 3)).astype(np.uint8)
            out.append(Image.fromarray(small).resize((fx["W"] * 8, fx["H"] * 8), Image.BICUBIC))
        return out

    return {"main": frames(1), "bg": frames(2), "objs": [frames(3 + j) for j in range(fx["n_obj"])]}


def loop_init_latents(fx) -> torch.Tensor:
    return torch.randn(1, 4, fx["T"], fx["H"], fx["W"], generator=torch.Generator().manual_seed(6))


def inversion_init_latents(fx) -> torch.Tensor:
    return torch.randn(1, 4, fx["T"], fx["H"], fx["W"], generator=torch.Generator().manual_seed(3000)) * 0.7


def source_latents(fx) -> list:
    """[bg, obj_1..obj_n] -> {t: [1,4,T,h,w]}: sqrt(a_t) x0 + sqrt(1-a_t) eps (stands in for inversion outputs)."""
    from mvoc_b200.scheduler import DDIMSchedule

    sched = DDIMSchedule(fx["n_steps"])
    shape = (1, 4, fx["T"], fx["H"], fx["W"])
    out = []
    for br in range(fx["n_obj"] + 1):
        x0 = torch.randn(shape, generator=torch.Generator().manual_seed(1000 + br))
        eps = torch.randn(shape, generator=torch.Generator().manual_seed(2000 + br))
        out.append({int(t): float(sched.alphas_cumprod[t]) ** 0.5 * x0 + (1 - float(sched.alphas_cumprod[t])) ** 0.5 * eps
                    for t in sched.timesteps})
    return out


def write_source_latents(root: str, fx) -> list:
    """The reference's wire format (pipeline_i2vgen_xl.py:1990-1993): one ddim_latents_{t}.pt per timestep."""
    import os

    dirs = []
    for br, per_t in enumerate(source_latents(fx)):
        d = os.path.join(root, f"branch{br}")
        os.makedirs(d, exist_ok=True)
        for t, x in per_t.items():
            torch.save(x.clone(), os.path.join(d, f"ddim_latents_{t}.pt"))
        dirs.append(d)
    return dirs


def write_mask_pngs(root: str) -> None:
    """Object 0: a folder of T frames (moving ellipse); object 1: one static PNG (rectangle).  8*W x 8*H pixels."""
    import os

    from PIL import Image, ImageDraw, ImageFilter

    os.makedirs(os.path.join(root, "obj0"), exist_ok=True)
    for i in range(T):
        im = Image.new("L", (W * 8, H * 8), 0)
        ImageDraw.Draw(im).ellipse([14 + 8 * i, 20, 54 + 8 * i, 64], fill=255)
        im.filter(ImageFilter.GaussianBlur(2.0)).save(os.path.join(root, "obj0", f"{i:03d}.png"))
    im = Image.new("L", (W * 8, H * 8), 0)
    ImageDraw.Draw(im).rectangle([70, 76, 118, 116], fill=230)
    im.filter(ImageFilter.GaussianBlur(1.5)).save(os.path.join(root, "obj1.png"))


def mask_paths(root: str) -> list:
    import os

    return [os.path.join(root, "obj0"), os.path.join(root, "obj1.png")]


def loop_inputs(fx, seam: dict) -> dict:
    """The `inputs` dict of oracle.pipeline.composite_loop / the product loop for a loop fixture: the seam tensors the
    reference's own conditioning code produced (stored in the golden file) + regenerated sources / noise + masks
    read from the committed PNGs by the product's mask_preprocess mirror."""
    import os

    from mvoc_b200.utils import mask_preprocess

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "masks128")
    masks = [mask_preprocess(p, "cpu", torch.float32, 1, 4, fx["T"]) for p in mask_paths(root)]
    return {"source_latents": source_latents(fx), "init_latents": loop_init_latents(fx),
            "prompt_embeds": seam["encoder_hidden_states"], "image_embeddings": seam["image_embeddings"],
            "image_latents_first": seam["image_latents_first"], "image_latents": seam["image_latents"],
            "fps": seam["fps"], "masks": masks}


def loop_workload(fx):
    from types import SimpleNamespace

    return SimpleNamespace(**{k: fx[k] for k in (
        "n_steps", "cfg", "n_obj", "pnp_f_t", "pnp_spatial_attn_t", "pnp_temp_attn_t", "inject_background",
        "fusion_step", "random_noise_ratio", "obj_random_noise_fusion", "obj_ddim_latents_idx_offset",
        "ddim_init_latents_t_idx")}, n_frames=fx["T"], latent_h=fx["H"], latent_w=fx["W"], n_branches=fx["n_obj"] + 3,
        unet="tiny4")
