"""`inverse.py` of the reference (i2vgen-xl/inverse.py) over the B200 path.

    python -m mvoc_b200.inverse --template_config configs/group_inversion/template.yaml \\
                                --configs_json configs/group_inversion/group_config.json

Same flags and config contract (inverse.py:230-255, :136-150).  Per active entry it runs the DDIM inversion
loop (pipelines/pipeline_i2vgen_xl.py:1940-2000; `inverse_config.n_steps`, cfg 1.0 => batch 1) and writes one
`ddim_latents_{t}.pt` per timestep into `inverse_config.output_dir` (:1988-1993) — the wire format
`composite.py` reads.  Like the reference it skips entries that were already inverted unless
`force_recompute_latents` (inverse.py:181-183) — "already inverted" means that EVERY expected
`ddim_latents_{t}.pt` is present (the reference only tests for the directory, so a crashed or `--max_steps` run is
never completed and composite.py later asserts on the missing timesteps).

Out of scope: frame loading and the VAE encode of inverse.py:49-55.  The clean video
latents / prompt / CLIP tensors are read from `--inputs file.pt` (dict: latents [1,4,T,h,w], prompt_embeds
[1,77,1024], image_embeddings [1,1,1024], image_latents [1,4,T,h,w]; `{video_name}` in the path selects one file
per entry).  A missing file is an error; `--synthetic` substitutes seeded synthetic tensors and random weights.
Independent videos are the unit of multi-GPU work here: with torchrun, entry i goes to rank i % world_size
(replicas, no collective).
"""
from __future__ import annotations

import argparse
import logging
import os

import torch

from . import config as cfgmod
from . import synthetic

logger = logging.getLogger(__name__)


def inversion_complete(out_dir: str, n_steps: int) -> bool:
    """True when every ddim_latents_{t}.pt of an n_steps inversion exists in out_dir."""
    from .pipeline import ddim_latents_filename
    from .scheduler import DDIMSchedule

    if not os.path.isdir(out_dir):
        return False
    return all(os.path.exists(os.path.join(out_dir, ddim_latents_filename(t)))
               for t in DDIMSchedule(n_steps, inverse=True).timesteps)


def main(template_config: str, configs_json: str, unet_state_dict: str = None, inputs: str = None,
         device: str = None, max_steps: int = None, synthetic_inputs: bool = False):
    from .pipeline import I2VGenXLPipeline
    from .unet3d import I2VGenXLUNet, UNetConfig, prepare

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    template = cfgmod.load_template(template_config)
    device = torch.device(device or (f"cuda:{local}" if world > 1 else template.get("device", "cuda:0")))
    if device.type == "cuda":
        torch.cuda.set_device(device)      # the C-ABI launches on the current device (mvoc_b200.ops._need_cuda)
    torch.set_grad_enabled(False)
    unet = I2VGenXLUNet(UNetConfig.full()).eval().requires_grad_(False)
    if unet_state_dict:
        unet.load_state_dict(torch.load(unet_state_dict, map_location="cpu"), strict=True)
    elif not synthetic_inputs:
        raise FileNotFoundError("--unet_state_dict is required (pass --synthetic to run on random-init weights)")
    unet = prepare(unet.to(device=device, dtype=torch.bfloat16))
    pipe = I2VGenXLPipeline(unet, device, use_cuda_graphs=True)
    done = 0
    for idx, config in enumerate(cfgmod.iter_configs(template_config, configs_json)):
        if idx % world != rank:
            continue                                                            # one video per GPU
        inv = config.inverse_config
        out_dir = inv.output_dir
        if inversion_complete(out_dir, int(inv.n_steps)) and not config.get("force_recompute_latents", False):
            logger.info("%s is complete: skipping (force_recompute_latents is false)", out_dir)   # inverse.py:181-183
            continue
        w_px, h_px = inv.image_size
        wl = synthetic.Workload(config.video_name, "full", int(inv.n_frames), int(h_px) // 8, int(w_px) // 8, 0,
                                inversion_steps=int(inv.n_steps))
        if inputs:
            ipath = inputs.format(video_name=config.video_name)
            if not os.path.exists(ipath):
                raise FileNotFoundError(f"inversion inputs {ipath!r} not found")
            src = torch.load(ipath, map_location="cpu")
        elif synthetic_inputs:
            src = synthetic.make_inversion_inputs(wl, idx)
        else:
            raise FileNotFoundError("no --inputs file (VAE / CLIP outputs are inputs of this script); pass "
                                    "--synthetic for seeded synthetic tensors")
        dt = unet.dtype
        pipe.invert(src["latents"].to(device).float().clone(), src["prompt_embeds"].to(device, dt),
                    src["image_embeddings"].to(device, dt), src["image_latents"].to(device, dt),
                    torch.full((1,), int(inv.get("target_fps", 8)), dtype=torch.int64, device=device),
                    num_inference_steps=int(inv.n_steps), guidance_scale=float(inv.cfg), output_dir=out_dir,
                    max_steps=max_steps, keep=False)
        done += 1
    return done


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--template_config", type=str, default="./configs/group_inversion/template.yaml")
    ap.add_argument("--configs_json", type=str, default="./configs/group_inversion/group_config.json")
    ap.add_argument("--unet_state_dict", type=str, default=None)
    ap.add_argument("--inputs", type=str, default=None)
    ap.add_argument("--device", type=str, default=None)
    ap.add_argument("--max_steps", type=int, default=None)
    ap.add_argument("--synthetic", action="store_true",
                    help="substitute seeded synthetic tensors for the inputs and random-init weights")
    return ap


if __name__ == "__main__":
    args = build_parser().parse_args()
    logging.basicConfig(level=logging.INFO)
    assert os.path.exists(args.template_config) and os.path.exists(args.configs_json)   # inverse.py:134
    main(args.template_config, args.configs_json, args.unet_state_dict, args.inputs, args.device, args.max_steps,
         args.synthetic)
