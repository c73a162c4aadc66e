"""Multi-GPU parity check (run under torchrun on the GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py [workload] [steps]

Every rank runs the frame-parallel composition loop; rank 0 also runs the single-GPU loop on the same inputs
and compares the final latents (only the order of the GroupNorm statistics merge differs between the two).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from mvoc_b200 import synthetic
    from mvoc_b200.parallel import FrameParallel
    from mvoc_b200.pipeline import Conditioning, I2VGenXLPipeline, LatentBank, init_pnp
    from mvoc_b200.scheduler import DDIMSchedule
    from mvoc_b200.unet3d import build_unet

    wl_name = sys.argv[1] if len(sys.argv) > 1 else "reduced2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    par = FrameParallel.from_env(dev)
    wl = synthetic.WORKLOADS[wl_name]
    sched = DDIMSchedule(wl.n_steps)
    inputs = synthetic.make_inputs(wl, sched.timesteps, sched.alphas_cumprod)
    unet = build_unet(wl.unet, seed=0, device=dev)
    bf = lambda x: x.to(device=dev, dtype=torch.bfloat16)

    def run(parallel, graphs=False):
        pipe = I2VGenXLPipeline(unet, dev, parallel=parallel, use_cuda_graphs=graphs)
        init_pnp(pipe, sched, wl)
        cond = Conditioning(bf(inputs["prompt_embeds"]), bf(inputs["image_embeddings"]),
                            bf(inputs["image_latents_first"]), bf(inputs["image_latents"]), inputs["fps"].to(dev))
        banks = [LatentBank(s, dev) for s in inputs["source_latents"]]
        masks = [(mf.to(dev), mb.to(dev)) for mf, mb in inputs["masks"]]
        lat = inputs["init_latents"].to(dev).clone()
        out = pipe.sample_with_pnp_pipeline_with_edit_prompt_extraction_with_attn_injection(
            cond, lat, banks[0], banks[1:], masks, num_inference_steps=wl.n_steps, guidance_scale=wl.cfg,
            fusion_steps=tuple(wl.fusion_step), random_noise_ratio=wl.random_noise_ratio, max_steps=steps)
        torch.cuda.synchronize()
        return out

    print(f"[multigpu_check] rank {par.rank}: {par.describe()}", flush=True)
    sharded = run(par)
    par.barrier()
    sharded_graph = run(par, graphs=True)     # the exchange kernels / NCCL collectives captured inside the CUDA graphs
    par.barrier()
    # the same partition with the exchange done by NCCL collectives: the peer-memory puts must move the same bytes
    peer, par.peer = par.peer, None
    sharded_nccl = run(par) if peer is not None else sharded
    par.peer = peer
    par.barrier()
    if par.rank == 0:
        single = run(FrameParallel.single(dev))
        err = float((sharded - single).norm() / single.norm())
        same = bool(torch.equal(sharded, sharded_graph))
        same_nccl = bool(torch.equal(sharded, sharded_nccl))
        print(f"[multigpu_check] {wl_name} x{par.world} ranks, {steps} steps: rel L2 vs single GPU = {err:.3e}; "
              f"graph replay == eager: {same}; peer-memory exchange == NCCL exchange: {same_nccl}")
        assert err <= 1e-2, err
        assert same and same_nccl
    par.barrier()
    par.shutdown()


if __name__ == "__main__":
    main()
