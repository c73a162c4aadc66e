"""Pins the oracle's MVOC layer (oracle/hooks.py, oracle/pipeline.py) against vectors produced by the
reference's OWN code (tests/golden/make_golden.py ran pnp_utils.register_*, composite.init_pnp and
I2VGenXLUnetExtension.forward from /root/reference).  CPU only."""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import hooks, ops_ref
from oracle import pipeline as opipe
from tests.golden import spec

GOLD = os.path.join(os.path.dirname(__file__), "golden", "unet_extension_forward_tiny4.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLD, map_location="cpu")


@pytest.mark.parametrize("case", spec.CASES, ids=[c["name"] for c in spec.CASES])
def test_oracle_reproduces_reference_unet_forward(golden, case):
    unet = spec.build_tiny4(seed=0)
    pipe = SimpleNamespace(unet=unet)
    wl = SimpleNamespace(n_steps=50, pnp_f_t=case["pnp_f_t"], pnp_spatial_attn_t=case["pnp_spatial_attn_t"],
                         pnp_temp_attn_t=case["pnp_temp_attn_t"], inject_background=case["inject_background"])
    opipe.init_pnp(pipe, torch.tensor(spec.timesteps_50()), wl)
    inp = spec.make_inputs(case)
    hooks.register_time_all(pipe, case["t"], inp["masks"])
    with torch.no_grad():
        y = opipe.unet_extension_forward(unet, inp["sample"], case["t"], inp["fps"], inp["image_latents_first"],
                                         inp["image_latents"], inp["image_embeddings"], inp["prompt_embeds"])
    ref = golden[case["name"]]
    err = float((y - ref).norm() / ref.norm())
    assert err <= 1e-5, f"{case['name']}: rel L2 {err:.3e} vs the reference's own output"


def test_golden_cases_are_distinct(golden):
    """Injection really changes the composite branches (and only them)."""
    a, b, c = golden["attn_only_t481"], golden["inject_bg_t481"], golden["no_hooks_t21"]
    assert not torch.equal(a[3:], b[3:])      # base = cond's own Q/K vs the background's
    assert torch.equal(a[:3], b[:3])          # source branches never depend on the composite ones
    assert not torch.equal(a, c)
    # uncond and cond composite branches share Q'/K' but keep their own V and context: different outputs
    assert not torch.equal(a[3], a[4])
    # on conv-injection steps conv_out's blend overwrites both composite outputs with the same tensor
    h = golden["all_hooks_t981"]
    assert torch.equal(h[3], h[4])


def test_injected_sites_match_reference_indices():
    """Layout rule == the reference's hard-coded res_dict {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}
    (pnp_utils.py:706, :889)."""
    unet = spec.build_tiny4(seed=0)
    assert hooks.injected_attention_sites(unet) == [(1, 1), (1, 2), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (3, 2)]


def test_wire_format_fixture():
    """ddim_latents_{t}.pt: one tensor [1,4,T,h,w] per timestep, name built from int(t) (utils.py:31-36)."""
    from mvoc_b200.pipeline import ddim_latents_filename, load_ddim_latents_at_t

    d = os.path.join(os.path.dirname(__file__), "golden", "ddim_latents_fixture")
    assert ddim_latents_filename(torch.tensor(981)) == "ddim_latents_981.pt"
    x = load_ddim_latents_at_t(981, d)
    assert x.shape == (1, 4, 4, 8, 8) and x.dtype == torch.float16


# ---------------------------------------------------------------- step loops run by the reference's own code
LOOPS = os.path.join(os.path.dirname(__file__), "golden", "composition_loop_tiny4.pt")
INVERSION = os.path.join(os.path.dirname(__file__), "golden", "inversion_loop_tiny4.pt")


LONG = os.environ.get("MVOC_LONG_TESTS") == "1"    # full 50-step / 500-step runs (2-3 minutes); default: first steps


@pytest.mark.parametrize("case,max_steps", [("default", 50 if LONG else 10), ("exotic", 10)])
def test_oracle_reproduces_reference_composition_loop(case, max_steps):
    """oracle.pipeline.composite_loop == the reference's sample_with_pnp_pipeline_with_edit_prompt_extraction_
    with_attn_injection (pipeline_i2vgen_xl.py:1552-1734, run by tests/golden/make_golden_loops.py): timestep and
    fusion-timestep selection, noise fusion, branch concat, hooks per step, CFG, DDIM update."""
    gold = torch.load(LOOPS, map_location="cpu")
    fx = spec.loop_fixture(case)
    rec = []
    opipe.composite_loop(spec.build_tiny4(seed=0), spec.loop_workload(fx), spec.loop_inputs(fx, gold["seam"]),
                         max_steps=max_steps, record=rec)
    checked = 0
    for i, ref in gold["latents_after_step"][case].items():
        if i < len(rec):
            err = float((rec[i] - ref).norm() / ref.norm())
            assert err <= 1e-5, f"{case} step {i}: {err:.3e}"
            checked += 1
    assert checked >= 5


def test_oracle_reproduces_reference_inversion_loop():
    """oracle.pipeline.invert_loop == the reference's invert (pipeline_i2vgen_xl.py:1914-2003); all 500 steps with MVOC_LONG_TESTS=1."""
    gold = torch.load(INVERSION, map_location="cpu")
    ix = spec.inversion_fixture()
    seam = gold["seam"]
    inv = {"latents": spec.inversion_init_latents(ix), "prompt_embeds": seam["encoder_hidden_states"],
           "image_embeddings": seam["image_embeddings"], "image_latents": seam["image_latents"], "fps": seam["fps"]}
    saved = opipe.invert_loop(spec.build_tiny4(seed=0), None, inv, n_steps=ix["n_steps"], max_steps=None if LONG else 3)
    assert len(saved) == (500 if LONG else 3)
    checked = 0
    for t, ref in gold["latents_at_t"].items():
        if t in saved:
            err = float((saved[t] - ref).norm() / ref.norm())
            assert err <= 1e-5, f"t={t}: {err:.3e}"
            checked += 1
    assert checked >= 3


def test_vae_decoder_restatement_structure():
    """oracle/vae.py: the full-width decoder has the parameter count of the published SD / i2vgen-xl VAE decoder
    (49 490 179 + the 4->4 post_quant_conv), decode_latents keeps the reference's layout (pipeline :771-791)."""
    from oracle import vae

    full = vae.build_decoder(narrow=False)
    assert sum(p.numel() for p in full.parameters()) == 49_490_179 + 20
    dec = vae.build_decoder()
    lat = torch.randn(2, 4, 3, 8, 8, generator=torch.Generator().manual_seed(0)) * vae.SCALING_FACTOR
    video = vae.decode_latents(dec, lat, decode_chunk_size=2)
    assert video.shape == (2, 3, 3, 64, 64) and video.dtype == torch.float32
    # frame f of video b depends only on latent frame f of video b
    lat2 = lat.clone()
    lat2[1, :, 2] += 0.1
    v2 = vae.decode_latents(dec, lat2)
    assert torch.equal(v2[0], video[0]) and torch.equal(v2[1, :, :2], video[1, :, :2])
    assert not torch.equal(v2[1, :, 2], video[1, :, 2])
    assert vae.psnr(video, video) == float("inf") and vae.psnr(v2, video) < 80
