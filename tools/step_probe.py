"""Is a composition step bound by the GPU or by the host enqueuing ~2700 launches?  (development aid)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mvoc_b200 import synthetic  # noqa: E402
from mvoc_b200.pipeline import Conditioning, I2VGenXLPipeline, LatentBank, init_pnp  # noqa: E402
from mvoc_b200.scheduler import DDIMSchedule  # noqa: E402
from mvoc_b200.unet3d import build_unet  # noqa: E402

dev = torch.device("cuda:0")
wl = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config2"]
sched = DDIMSchedule(wl.n_steps)
inputs = synthetic.make_inputs(wl, sched.timesteps, sched.alphas_cumprod)
torch.backends.cudnn.benchmark = True
unet = build_unet(wl.unet, seed=0, device=dev)
pipe = I2VGenXLPipeline(unet, dev)
init_pnp(pipe, sched, wl)
bf = lambda x: x.to(device=dev, dtype=torch.bfloat16)
cond = Conditioning(bf(inputs["prompt_embeds"]), bf(inputs["image_embeddings"]), bf(inputs["image_latents_first"]),
                    bf(inputs["image_latents"]), inputs["fps"].to(dev))
banks = [LatentBank(s, dev) for s in inputs["source_latents"]]
masks = [(mf.to(dev), mb.to(dev)) for mf, mb in inputs["masks"]]


def run(start, n):
    lat = inputs["init_latents"].to(dev).clone()
    return pipe.sample_with_pnp_pipeline_with_edit_prompt_extraction_with_attn_injection(
        cond, lat, banks[0], banks[1:], masks, num_inference_steps=wl.n_steps, guidance_scale=wl.cfg,
        fusion_steps=tuple(wl.fusion_step), random_noise_ratio=wl.random_noise_ratio, start_step=start, max_steps=n)


run(0, 8)
torch.cuda.synchronize()
for start in (0, 6):
    t0 = time.perf_counter()
    run(start, 4)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"steps {start}..{start+3}: host enqueue {1e3*(t1-t0)/4:.1f} ms/step, total {1e3*(t2-t0)/4:.1f} ms/step")
