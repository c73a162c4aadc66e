"""Multi-GPU partition of the composition path: one process per GPU, frames sharded across ranks.

The path shards naturally along the frame axis: resnets, spatial transformers,
cross-attention, down/up-sampling and every mask blend are independent per frame, so each rank owns
T/P frames of ALL n_obj+3 branches — which also keeps MVOC's source->composite Q/K injection local to a
rank (no Q/K exchange at all).  Only the temporal operators (temporal conv, temporal attention,
5-D GroupNorm statistics) need every frame of a pixel; around those the activation is re-laid out from
frame shards to pixel shards and back with one NCCL all-to-all each way over NVLink/NVSwitch.

world_size == 1 is a strict no-op (no process group is created).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


class FrameParallel:
    def __init__(self, group=None, world: int = 1, rank: int = 0, device=None):
        self.group = group
        self.world = world
        self.rank = rank
        self.device = device

    # ------------------------------------------------------------------ setup
    @classmethod
    def single(cls, device=None) -> "FrameParallel":
        return cls(None, 1, 0, device)

    @classmethod
    def from_env(cls, device=None, backend: Optional[str] = None) -> "FrameParallel":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world == 1:
            return cls.single(device)
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if (device is not None and torch.device(device).type == "cuda") else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend=backend)
        return cls(dist.group.WORLD, dist.get_world_size(), dist.get_rank(), device)

    def shutdown(self) -> None:
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()

    def describe(self) -> str:
        return "single GPU" if self.world == 1 else f"frame-parallel x{self.world} (all-to-all around temporal ops)"

    # ------------------------------------------------------------------ plumbing
    def barrier(self) -> None:
        if self.world > 1:
            dist.barrier(group=self.group)

    def max_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    # ------------------------------------------------------------------ shard geometry
    def frame_range(self, n_frames: int):
        """(lo, hi) of the frames this rank owns."""
        if n_frames % self.world != 0:
            raise ValueError(f"n_frames={n_frames} is not divisible by world_size={self.world}")
        per = n_frames // self.world
        return (self.rank * per, (self.rank + 1) * per)

    def pixel_range(self, n_pixels: int):
        """(lo, hi) of the pixels this rank owns while a temporal operator runs."""
        if n_pixels % self.world != 0:
            raise ValueError(f"h*w={n_pixels} is not divisible by world_size={self.world}")
        per = n_pixels // self.world
        return (self.rank * per, (self.rank + 1) * per)

    # ------------------------------------------------------------------ re-layout (the exchange step)
    def to_pixel_shards(self, x: torch.Tensor, num_frames: int) -> torch.Tensor:
        """[(b t_local), h, w, C] (this rank's frames, all pixels) -> [(b T), 1, S/P, C] (all frames, this
        rank's pixels): one all-to-all.  Rows stay in (video, frame, pixel) order."""
        P = self.world
        btl, h, w, C = x.shape
        tl = num_frames // P
        b, S = btl // tl, h * w
        sp = S // P
        if S % P != 0:
            raise ValueError(f"h*w={S} is not divisible by world_size={P}")
        send = x.contiguous().view(b, tl, P, sp, C).permute(2, 0, 1, 3, 4).contiguous()   # [dst, b, tl, sp, C]
        recv = torch.empty_like(send)                                            # [src, b, tl, sp, C]
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.permute(1, 0, 2, 3, 4).contiguous().view(b * num_frames, 1, sp, C)   # frame = src*tl + i

    def to_frame_shards(self, y: torch.Tensor, num_frames: int, h: int, w: int) -> torch.Tensor:
        """Inverse of to_pixel_shards: [(b T), 1, S/P, C] -> [(b t_local), h, w, C]."""
        P = self.world
        bT, _, sp, C = y.shape
        b, tl = bT // num_frames, num_frames // P
        send = y.contiguous().view(b, P, tl, sp, C).permute(1, 0, 2, 3, 4).contiguous()   # [dst(frame owner), b, tl, sp, C]
        recv = torch.empty_like(send)                                            # [src(pixel owner), b, tl, sp, C]
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.permute(1, 2, 0, 3, 4).contiguous().view(b * tl, h, w, C)    # pixel = src*sp + j

    def gather_partials(self, partial: torch.Tensor) -> torch.Tensor:
        """GroupNorm partial statistics of every rank's pixel shard: [N,G,chunks,2] -> [P,N,G,chunks,2]."""
        partial = partial.contiguous()
        out = torch.empty((self.world * partial.shape[0],) + tuple(partial.shape[1:]), dtype=partial.dtype,
                          device=partial.device)
        dist.all_gather_into_tensor(out, partial, group=self.group)   # concatenated along dim 0
        return out.view((self.world,) + tuple(partial.shape))

    def gather_frames(self, x: torch.Tensor) -> torch.Tensor:
        """[k, t_local, ...] on every rank -> [k, T, ...] (frames in rank order)."""
        x = x.contiguous()
        out = torch.empty((self.world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=self.group)
        out = out.view((self.world,) + tuple(x.shape))
        k, tl = x.shape[0], x.shape[1]
        return out.permute(1, 0, 2, *range(3, out.dim())).contiguous().view(k, self.world * tl, *x.shape[2:])

    # ------------------------------------------------------------------ temporal operators, replicated
    # When a level's h*w does not divide by the number of ranks (config 5's 11x20 level on 8 GPUs) — or is at
    # most MVOC_FP_GATHER_MAX_PIXELS (off by default; the low-resolution levels are latency-bound) — the frames
    # are all-gathered instead, every rank runs the temporal operator on all pixels and keeps its own frames:
    # one collective instead of two, no GroupNorm statistics exchange, P-fold redundant work on a small tensor.
    _GATHER_MAX_PIXELS = int(os.environ.get("MVOC_FP_GATHER_MAX_PIXELS", "0") or 0)

    def _replicate(self, h: int, w: int) -> bool:
        return (h * w) % self.world != 0 or h * w <= self._GATHER_MAX_PIXELS

    def _temporal_replicated(self, module, hidden_states: torch.Tensor, num_frames: int) -> torch.Tensor:
        btl, h, w, C = hidden_states.shape
        tl = num_frames // self.world
        b = btl // tl
        full = self.gather_frames(hidden_states.contiguous().view(b, tl, h, w, C)).view(b * num_frames, h, w, C)
        y = module.forward_local(full, num_frames, None)
        f0, f1 = self.frame_range(num_frames)
        return y.view(b, num_frames, h, w, C)[:, f0:f1].contiguous().view(b * tl, h, w, C)

    # ------------------------------------------------------------------ temporal operators on pixel shards
    def temporal_transformer(self, module, hidden_states: torch.Tensor, num_frames: int) -> torch.Tensor:
        _, h, w, _ = hidden_states.shape
        if self._replicate(h, w):
            return self._temporal_replicated(module, hidden_states, num_frames)   # ctx.full_hw stays None
        xs = self.to_pixel_shards(hidden_states, num_frames)
        module.ctx.full_hw = (h, w)
        try:
            ys = module.forward_local(xs, num_frames, self.gather_partials)
        finally:
            module.ctx.full_hw = None
        return self.to_frame_shards(ys, num_frames, h, w)

    def temporal_conv(self, module, hidden_states: torch.Tensor, num_frames: int) -> torch.Tensor:
        _, h, w, _ = hidden_states.shape
        if self._replicate(h, w):
            return self._temporal_replicated(module, hidden_states, num_frames)
        xs = self.to_pixel_shards(hidden_states, num_frames)
        ys = module.forward_local(xs, num_frames, self.gather_partials)
        return self.to_frame_shards(ys, num_frames, h, w)
