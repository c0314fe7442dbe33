"""Data-parallel plumbing: one process per GPU, trajectories sharded over ranks, ONE gradient all-reduce per
step over NCCL/NVLink.  The reference has no DistributedDataParallel; its only helper is
libs/pino_utils/distributed.py:23-34 (all_reduce then divide by world size) -- same semantics here, applied
to a flat gradient bucket (complex grads viewed as real)."""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of independent trajectories over ranks."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class GradBucket:
    """Flat fp32 bucket over all parameter gradients; `.allreduce_mean()` sums over ranks and divides by
    the world size, then scatters the views back (grads become views into the bucket after the first call,
    so later steps all-reduce in place with no copies)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.offsets: List[int] = []
        off = 0
        for p in self.params:                      # 16-byte aligned slots (view_as_complex needs even offsets)
            self.offsets.append(off)
            off += -(-((2 if p.is_complex() else 1) * p.numel()) // 4) * 4
        self.numel = off
        self.flat: Optional[torch.Tensor] = None

    def _real_view(self, t: torch.Tensor) -> torch.Tensor:
        return torch.view_as_real(t) if t.is_complex() else t

    def attach(self):
        """Make every .grad a view into the flat bucket (zero-initialised)."""
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            n = (2 if p.is_complex() else 1) * p.numel()
            chunk = self.flat[off: off + n]
            if p.is_complex():
                view = torch.view_as_complex(chunk.view(*p.shape, 2))
            else:
                view = chunk.view(p.shape)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view

    def zero(self):
        if self.flat is not None:
            self.flat.zero_()

    def allreduce_mean(self, group=None):
        if self.flat is None:
            self.attach()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat


def all_reduce_mean_scalar(t: torch.Tensor, group=None) -> torch.Tensor:
    """libs/pino_utils/distributed.py:23-34."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t / dist.get_world_size(group)
    return t
