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

    def _views(self):
        if getattr(self, "_view_cache", None) is None or self._view_cache[0] is not self.flat:
            views = []
            for p, off in zip(self.params, self.offsets):
                n = (2 if p.is_complex() else 1) * p.numel()
                chunk = self.flat[off: off + n]
                views.append(torch.view_as_complex(chunk.view(*p.shape, 2)) if p.is_complex() else chunk.view(p.shape))
            self._view_cache = (self.flat, views)
        return self._view_cache[1]

    def release_grads(self):
        """p.grad = None for every parameter: the next backward then ASSIGNS its gradient tensors (no `grad += g`
        kernel per parameter); `gather()` collects them into the flat bucket with one launch."""
        for p in self.params:
            p.grad = None

    def gather(self):
        """Copy whatever gradient tensors the backward left in p.grad into the flat bucket (ONE kernel over a device
        pointer table; missing gradients become zeros) and make p.grad views of the bucket again.  CUDA only.
        Inside a CUDA-graph capture the pointer table is captured with the graph (the gradient tensors live in the
        graph's private pool, so their addresses are the same at every replay)."""
        from . import ops
        if self.flat is None:
            self.attach()
            return self.flat
        views = self._views()
        nseg = len(self.params)
        dev = self.flat.device
        if getattr(self, "_seg_dev", None) is None:
            offs = list(self.offsets) + [self.numel]
            cnts = [(2 if p.is_complex() else 1) * p.numel() for p in self.params]
            self._seg_dev = (torch.tensor(offs, dtype=torch.int64, device=dev), torch.tensor(cnts, dtype=torch.int64, device=dev))
            self._ptr_dev = torch.zeros(nseg, dtype=torch.int64, device=dev)
            # pinned host pointer tables, allocated ONCE (no cudaHostAlloc inside a CUDA-graph capture); two of them so
            # that a table a captured graph copies from at every replay is never rewritten by a later eager call
            self._ptr_host = [torch.zeros(nseg, dtype=torch.int64).pin_memory() for _ in range(2)]
            self._keep = []
        ptrs, keep, in_place = [], [], True
        for p, v in zip(self.params, views):
            g = p.grad
            if g is None:
                ptrs.append(0)
                in_place = False
                continue
            if g.device != dev or g.dtype != p.dtype:
                raise RuntimeError("GradBucket.gather: gradient on the wrong device / of the wrong dtype")
            if not g.is_contiguous():
                g = g.contiguous()
            keep.append(g)
            ptrs.append(g.data_ptr())
            in_place = in_place and g.data_ptr() == v.data_ptr()
        if not in_place:
            capturing = torch.cuda.is_current_stream_capturing()
            # the host table must stay untouched for as long as a captured graph may replay the copy: captures own table 1
            host = self._ptr_host[1 if capturing else 0]
            if not capturing and getattr(self, "_host_copied", None) is not None:
                self._host_copied.synchronize()          # the previous call's async copy out of this table has finished
            host.copy_(torch.tensor(ptrs, dtype=torch.int64))
            if capturing:
                self._keep.extend(keep)
            self._ptr_dev.copy_(host, non_blocking=True)
            if not capturing:
                self._host_copied = torch.cuda.Event()
                self._host_copied.record(torch.cuda.current_stream(dev))
            ops.gather_segments(self.flat, self._ptr_dev, self._seg_dev[0], self._seg_dev[1], nseg)
            if not capturing:
                # eager: the source tensors / host table may be freed once the copy + kernel are enqueued only because
                # the caching allocators are stream-ordered; keep them until the next call to be explicit
                self._last = keep
        for p, v in zip(self.params, views):
            p.grad = v
        return self.flat

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
