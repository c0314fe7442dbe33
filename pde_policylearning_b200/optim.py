"""Fused flat-buffer Adam and the CUDA-graph training step.

The reference trains with ``torch.optim.Adam(model.parameters(), lr, weight_decay)`` (run_pde_observers.py:134,
train_pino.py:205): ~30 tiny foreach kernels per step for a 600 k-parameter FNO2d.  Here every parameter, gradient
and moment lives in ONE flat fp32 buffer (complex parameters as their real view, which is how torch.optim.Adam
treats them), the update is one kernel (csrc/optim.cu), and the 1/world_size of the data-parallel mean is folded
into it.  ``GraphedTrainStep`` captures forward + loss + backward + all-reduce + Adam into one CUDA graph."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from .parallel import GradBucket


class FusedAdam:
    """Adam with torch.optim.Adam's update rule over a flat parameter buffer.

    After construction every ``p.data`` is a view into ``self.flat_param`` and every ``p.grad`` a view into
    ``self.bucket.flat`` (so ``state_dict()`` / checkpoints of the MODEL are unchanged)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 bucket: Optional[GradBucket] = None, overlap_allreduce: bool = False, bucket_bytes: int = 32 << 20,
                 group=None):
        """overlap_allreduce: sum the gradients over the ranks bucket by bucket DURING the backward (parallel.OverlappedGradSync)
        instead of with one all-reduce after it."""
        self.bucket = bucket if bucket is not None else GradBucket(params)
        self._overlap_cfg = (bool(overlap_allreduce), int(bucket_bytes), group)
        self.overlap = None
        ps = self.bucket.params
        if not ps:
            raise ValueError("FusedAdam got no trainable parameters")
        dev = ps[0].device
        ops._require_cuda(*ps)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        n = self.bucket.numel
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, off in zip(ps, self.bucket.offsets):
            k = (2 if p.is_complex() else 1) * p.numel()
            chunk = self.flat_param[off: off + k]
            view = torch.view_as_complex(chunk.view(*p.shape, 2)) if p.is_complex() else chunk.view(p.shape)
            view.copy_(p.data)
            p.data = view
        if self.bucket.flat is None:
            self.bucket.attach()
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=dev)
        if self._overlap_cfg[0]:
            from .parallel import OverlappedGradSync
            self.overlap = OverlappedGradSync(self.bucket, self._overlap_cfg[1], self._overlap_cfg[2])

    def zero_grad(self, set_to_none: bool = False):
        """set_to_none: drop the gradients, so that the next backward assigns them and `sync_grads()` collects
        them with one kernel; False: zero the flat bucket in place (gradients then accumulate into it, as torch does)."""
        if set_to_none:
            self.bucket.release_grads()
        else:
            for p, v in zip(self.bucket.params, self.bucket._views()):
                p.grad = v
            self.bucket.zero()

    def sync_grads(self, group=None):
        """After backward: gather the per-parameter gradients into the flat bucket and (world > 1) sum them over the
        ranks with ONE all-reduce; `step(grad_scale=1/world)` turns the sum into the data-parallel mean."""
        if self.overlap is not None:
            self.overlap.finish()          # the per-bucket all-reduces were launched from the backward's gradient hooks
            return
        self.bucket.gather()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=group)

    def step(self, grad_scale: float = 1.0):
        self.bucket.gather()           # no-op when the gradients already live in the flat bucket
        L = _lib.lib()
        _lib.check(L.b2no_adam_step(ops._ptr(self.flat_param), ops._ptr(self.bucket.flat), ops._ptr(self.exp_avg),
                                    ops._ptr(self.exp_avg_sq), self.bucket.numel, ops._ptr(self.step_counter),
                                    self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                    float(grad_scale), ops._stream()), "adam_step")

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_counter,
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_counter.copy_(sd["step"])
        self.lr, self.betas, self.eps, self.weight_decay = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["weight_decay"]


class GraphedTrainStep:
    """One training step -- zero grads, forward, loss, backward, (sum all-reduce), fused Adam -- captured ONCE
    into a CUDA graph and replayed: the ~150 launches of a step cost one cudaGraphLaunch on the host.

        step = GraphedTrainStep(model, loss_fn, opt, example_inputs, example_target)
        loss = step(inputs, target)          # copies into the static buffers, replays, returns the loss tensor

    ``loss_fn(output, target)``; inputs is a tuple of tensors (device or pinned host: the copy into the static
    input buffers is the H2D transfer)."""

    def __init__(self, model, loss_fn: Callable, opt: FusedAdam, example_inputs, example_target, warmup: int = 3,
                 group=None):
        self.model, self.loss_fn, self.opt, self.group = model, loss_fn, opt, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = opt.flat_param.device
        self.static_in = [torch.empty_like(t, device=dev).copy_(t) for t in example_inputs]
        self.static_tgt = torch.empty_like(example_target, device=dev).copy_(example_target)
        # warm-up on a side stream (plans, tensor maps, allocator pools), then restore the optimizer state so that
        # construction does not advance training
        saved = (opt.flat_param.clone(), opt.exp_avg.clone(), opt.exp_avg_sq.clone(), opt.step_counter.clone())
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager()
        self.launches_per_step = ops.launch_count() - l0
        for dst, src in zip((opt.flat_param, opt.exp_avg, opt.exp_avg_sq, opt.step_counter), saved):
            dst.copy_(src)

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        out = self.model(*self.static_in)
        loss = self.loss_fn(out, self.static_tgt)
        loss.backward()
        self.opt.sync_grads(self.group)
        self.opt.step(grad_scale=1.0 / self.world)
        return loss.detach()

    def __call__(self, inputs, target):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                if tuple(src.shape) != tuple(dst.shape):      # copy_ would broadcast a ragged last batch silently
                    raise ValueError(f"GraphedTrainStep: input of shape {tuple(src.shape)}, the graph was captured for "
                                     f"{tuple(dst.shape)}")
                dst.copy_(src, non_blocking=True)
        if self.static_tgt.data_ptr() != target.data_ptr():
            if tuple(target.shape) != tuple(self.static_tgt.shape):
                raise ValueError(f"GraphedTrainStep: target of shape {tuple(target.shape)}, the graph was captured for "
                                 f"{tuple(self.static_tgt.shape)}")
            self.static_tgt.copy_(target, non_blocking=True)
        self.graph.replay()
        ops._REPLAYED[0] += self.launches_per_step
        return self.static_loss

    def close(self):
        """Release the captured graph (call before torch.distributed.destroy_process_group(): destroying the NCCL
        communicator while a graph that captured its all-reduce is alive blocks)."""
        if self.graph is not None:
            self.graph.reset()
            self.graph = None


class HostBatchPipeline:
    """Feeds pinned HOST batches to a GraphedTrainStep with the host->device copy of batch i+1 running on a copy stream
    while step i computes (two device staging sets, events both ways).  Every batch is still copied host->device once;
    only its latency leaves the critical path.

        pipe = HostBatchPipeline(step)
        for loss in pipe.run(batches):       # batches: iterable of (inputs_tuple, target), pinned host tensors
            value = loss.item()
    """

    def __init__(self, step: GraphedTrainStep):
        self.step = step
        dev = step.static_tgt.device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [([torch.empty_like(t) for t in step.static_in], torch.empty_like(step.static_tgt)) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]      # staging set filled
        self.free = [torch.cuda.Event() for _ in range(2)]       # staging set consumed by the step

    def _upload(self, slot, inputs, target, first_use):
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.free[slot])
            else:
                self.copy_stream.wait_stream(main)
            ins, tgt = self.stage[slot]
            for dst, src in zip(ins, inputs):
                dst.copy_(src, non_blocking=True)
            tgt.copy_(target, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def run_losses(self, batches):
        """Like run(), but yields the loss of every step as a Python float read back through a pinned host buffer ONE
        STEP LATE: the 4-byte device->host copy of step i is enqueued right after its replay and awaited only after step
        i+1 has been launched, so the GPU never idles on the host's read-back (every step's loss is still read)."""
        dev = self.step.static_tgt.device
        host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        main = torch.cuda.current_stream(dev)
        i = 0
        for loss in self.run(batches):
            host[i & 1].copy_(loss, non_blocking=True)
            done[i & 1].record(main)
            if i > 0:
                done[(i - 1) & 1].synchronize()
                yield float(host[(i - 1) & 1])
            i += 1
        if i > 0:
            done[(i - 1) & 1].synchronize()
            yield float(host[(i - 1) & 1])

    def run(self, batches):
        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        main = torch.cuda.current_stream()
        self._upload(0, nxt[0], nxt[1], True)
        i = 0
        while nxt is not None:
            slot = i & 1
            try:
                after = next(it)
            except StopIteration:
                after = None
            if after is not None:
                self._upload(slot ^ 1, after[0], after[1], i == 0)
            main.wait_event(self.ready[slot])
            ins, tgt = self.stage[slot]
            loss = self.step(tuple(ins), tgt)        # device->device into the graph's static buffers, then replay
            self.free[slot].record(main)
            yield loss
            nxt = after
            i += 1

