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

    # ------------------------------------------------------------------ frame sharding
    def frame_slice(self, n_frames: int) -> slice:
        if n_frames % self.world != 0:
            raise ValueError(f"n_frames={n_frames} is not divisible by world_size={self.world}")
        per = n_frames // self.world
        return slice(self.rank * per, (self.rank + 1) * per)

    def local_frames(self, n_frames: int) -> int:
        return n_frames // self.world if self.world > 1 else n_frames
