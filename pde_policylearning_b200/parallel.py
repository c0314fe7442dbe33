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


class OverlappedGradSync:
    """Gradient all-reduce launched DURING the backward, bucket by bucket ("bucket per layer, launch when that
    layer's dW is complete, overlap with the remaining backward").

    The parameters are split, from the LAST one backwards (the order in which the backward finishes them), into groups
    of about `bucket_bytes`; every group is one contiguous range of the flat bucket.  A post-accumulate hook on every
    parameter counts arrivals; when a group is complete its gradients are gathered into the flat range (one kernel) and
    the range is summed over the ranks with ONE NCCL all-reduce issued on a side stream, so the collective of the late
    layers runs underneath the backward of the early ones.  `finish()` (called by FusedAdam.sync_grads) launches what is
    left, makes the compute stream wait for the side stream, and points every p.grad at its bucket view.  Works inside a
    CUDA-graph capture (the side stream is forked from and joined to the capturing stream).  Semantics are those of
    libs/pino_utils/distributed.py:23-34 (sum, then divide by the world size -- the division is folded into Adam)."""

    def __init__(self, bucket: GradBucket, bucket_bytes: int = 32 << 20, group=None):
        self.bucket, self.group = bucket, group
        ps = bucket.params
        if bucket.flat is None:
            bucket.attach()
        n = len(ps)
        ends = [bucket.offsets[i + 1] if i + 1 < n else bucket.numel for i in range(n)]
        self.ranges = []                       # (lo, hi): parameters lo .. hi-1, in launch order (last parameters first)
        hi, acc = n, 0
        for i in range(n - 1, -1, -1):
            acc += (ends[i] - bucket.offsets[i]) * 4
            if acc >= bucket_bytes or i == 0:
                self.ranges.append((i, hi))
                hi, acc = i, 0
        self.group_of = [0] * n
        for g, (lo, hi) in enumerate(self.ranges):
            for i in range(lo, hi):
                self.group_of[i] = g
        self.flat_range = [(bucket.offsets[lo], ends[hi - 1]) for lo, hi in self.ranges]
        self._reset()
        self.enabled = True
        self.comm_stream = torch.cuda.Stream(device=ps[0].device) if ps[0].is_cuda else None
        self._host_tables = None
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(ps)]
        self.launched_during_backward = 0      # evidence for tests / bench: groups whose all-reduce started before finish()

    def _reset(self):
        self._pending = [hi - lo for lo, hi in self.ranges]
        self._launched = [False] * len(self.ranges)

    def _make_hook(self, i):
        def hook(p):
            if not self.enabled:
                return
            g = self.group_of[i]
            self._pending[g] -= 1
            if self._pending[g] == 0 and not self._launched[g]:
                self._launch(g)
                self.launched_during_backward += 1
        return hook

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _gather_group(self, g):
        b = self.bucket
        lo, hi = self.ranges[g]
        views = b._views()
        if b.flat.is_cuda:
            from . import ops
            if getattr(b, "_seg_dev", None) is None:
                dev = b.flat.device
                offs = list(b.offsets) + [b.numel]
                cnts = [(2 if p.is_complex() else 1) * p.numel() for p in b.params]
                b._seg_dev = (torch.tensor(offs, dtype=torch.int64, device=dev), torch.tensor(cnts, dtype=torch.int64, device=dev))
                b._ptr_dev = torch.zeros(len(b.params), dtype=torch.int64, device=dev)
                b._ptr_host = [torch.zeros(len(b.params), dtype=torch.int64).pin_memory() for _ in range(2)]
                b._keep = []
            capturing = torch.cuda.is_current_stream_capturing()
            host = b._ptr_host[1 if capturing else 0]
            keep = []
            for i in range(lo, hi):
                gr = b.params[i].grad
                if gr is not None and not gr.is_contiguous():
                    gr = gr.contiguous()
                keep.append(gr)
                host[i] = 0 if gr is None else gr.data_ptr()
            (b._keep if capturing else self._keep_eager).extend(keep)
            b._ptr_dev[lo:hi].copy_(host[lo:hi], non_blocking=True)
            # offsets are absolute positions in the flat bucket, so the sub-tables address the same destination
            ops.gather_segments(b.flat, b._ptr_dev[lo:hi], b._seg_dev[0][lo:hi + 1], b._seg_dev[1][lo:hi], hi - lo)
        else:
            for i in range(lo, hi):
                gr = b.params[i].grad
                if gr is None:
                    views[i].zero_()
                elif gr.data_ptr() != views[i].data_ptr():
                    views[i].copy_(gr)

    def _launch(self, g):
        self._gather_group(g)
        self._launched[g] = True
        if self._world() > 1:
            fs, fe = self.flat_range[g]
            chunk = self.bucket.flat[fs:fe]
            if self.comm_stream is not None:
                cur = torch.cuda.current_stream(self.bucket.flat.device)
                self.comm_stream.wait_stream(cur)
                with torch.cuda.stream(self.comm_stream):
                    dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)

    _keep_eager: List = []

    def finish(self):
        """Launch the groups the backward did not complete (parameters without a gradient count as zeros), join the side
        stream, re-point p.grad at the bucket views."""
        self._keep_eager = self._keep_eager[-4 * len(self.bucket.params):]
        for g in range(len(self.ranges)):
            if not self._launched[g]:
                self._launch(g)
        if self.comm_stream is not None and self._world() > 1:
            torch.cuda.current_stream(self.bucket.flat.device).wait_stream(self.comm_stream)
        was = self.enabled
        self.enabled = False
        for p, v in zip(self.bucket.params, self.bucket._views()):
            p.grad = v
        self.enabled = was
        self._reset()
        return self.bucket.flat

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def all_reduce_mean_scalar(t: torch.Tensor, group=None) -> torch.Tensor:
    """libs/pino_utils/distributed.py:23-34."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t = t / dist.get_world_size(group)
    return t
