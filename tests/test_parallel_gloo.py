"""world_size-2 (and 4) gloo tests of the multi-GPU host logic on CPU: the frame<->pixel re-layout
all-to-all, the gathers, shard geometry and the sharded mask tokens."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mvoc_b200.parallel import FrameParallel

        par = FrameParallel(dist.group.WORLD, world, rank, torch.device("cpu"))
        globals()[fn_name](par)
        ret[rank] = "ok"
    except Exception as ex:  # surfaced in the parent
        import traceback

        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def _run(world, fn_name):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", f"rank {r}: {ret.get(r)}"


def _full_tensor(b, T, h, w, C):
    return torch.arange(b * T * h * w * C, dtype=torch.float32).view(b, T, h, w, C)


def check_relayout_roundtrip(par):
    # second shape: h == world and S/world == w, where a permute+reshape can silently stay a strided VIEW
    # (the 8-GPU run at the 8x8 level hit exactly that) — outputs must be contiguous
    for (b, T, h, w, C) in [(3, 8, 4, 6, 5), (2, 2 * par.world, par.world, 3, 4)]:
        _relayout_case(par, b, T, h, w, C)


def _relayout_case(par, b, T, h, w, C):
    full = _full_tensor(b, T, h, w, C)
    f0, f1 = par.frame_range(T)
    local = full[:, f0:f1].reshape(b * (f1 - f0), h, w, C).contiguous()
    px = par.to_pixel_shards(local, T)
    p0, p1 = par.pixel_range(h * w)
    expect = full.view(b, T, h * w, C)[:, :, p0:p1].reshape(b * T, 1, p1 - p0, C)
    assert torch.equal(px, expect), "pixel shard does not hold (all frames) x (own pixels) in (b, t, p) order"
    back = par.to_frame_shards(px, T, h, w)
    assert torch.equal(back, local)
    assert px.is_contiguous() and back.is_contiguous()


def check_gathers(par):
    part = torch.full((4, 2, 3, 2), float(par.rank))
    allp = par.gather_partials(part)
    assert allp.shape == (par.world, 4, 2, 3, 2)
    for r in range(par.world):
        assert torch.all(allp[r] == r)
    k, T, h = 2, 8, 3
    full = torch.arange(k * T * h, dtype=torch.float32).view(k, T, h)
    f0, f1 = par.frame_range(T)
    got = par.gather_frames(full[:, f0:f1].contiguous())
    assert torch.equal(got, full)
    assert par.max_over_ranks(float(par.rank)) == float(par.world - 1)
    par.barrier()


def check_sharded_mask_tokens(par):
    """Every rank's token masks are the matching slices of the single-GPU ones."""
    from mvoc_b200 import pnp_utils
    from mvoc_b200.synthetic import make_masks

    T, H, W, h, w = 8, 16, 16, 8, 8
    masks = make_masks(2, T, H, W, seed=2)
    cache = pnp_utils._MaskCache()
    full_bin = cache.tokens(masks, h, w, soft=False).view(2, T, h * w)
    full_soft = cache.tokens(masks, h, w, soft=True).view(2, T, h * w)
    f0, f1 = par.frame_range(T)
    p0, p1 = par.pixel_range(h * w)
    assert torch.equal(cache.tokens(masks, h, w, soft=False, frames=(f0, f1)).view(2, f1 - f0, h * w), full_bin[:, f0:f1])
    assert torch.equal(cache.tokens(masks, h, w, soft=True, pixels=(p0, p1)).view(2, T, p1 - p0), full_soft[:, :, p0:p1])


@pytest.mark.parametrize("world", [2, 4])
def test_frame_pixel_relayout(world):
    _run(world, "check_relayout_roundtrip")


def test_gathers_world2():
    _run(2, "check_gathers")


def test_sharded_mask_tokens_world2():
    _run(2, "check_sharded_mask_tokens")


def test_single_rank_is_a_noop():
    sys.path.insert(0, ROOT)
    from mvoc_b200.parallel import FrameParallel

    par = FrameParallel.single()
    assert par.world == 1 and par.frame_range(16) == (0, 16) and par.max_over_ranks(3.5) == 3.5
    par.barrier()
    assert "single" in par.describe()
