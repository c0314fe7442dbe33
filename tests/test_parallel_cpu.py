"""CPU, gloo, world_size 2: the data-parallel host logic (sharding + flat-bucket gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pde_policylearning_b200 import parallel
    r, lr, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    cw = torch.nn.Parameter(torch.randn(2, 2, dtype=torch.cfloat))
    params = list(lin.parameters()) + [cw]
    bucket = parallel.GradBucket(params)
    bucket.attach()
    data = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4)
    mine = data[list(parallel.shard_range(8, rank, world))]
    loss = lin(mine).sum() + (cw * (rank + 1)).abs().sum()
    loss.backward()
    bucket.allreduce_mean()
    # grads are views into the bucket and identical on all ranks
    gathered = [torch.zeros_like(bucket.flat) for _ in range(world)]
    dist.all_gather(gathered, bucket.flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # expected: mean over ranks of the per-rank gradient
    exp_w = data.reshape(world, -1, 4).sum(dim=1).mean(dim=0)
    assert torch.allclose(lin.weight.grad[0], exp_w)
    assert cw.grad.data_ptr() >= bucket.flat.data_ptr()
    s = parallel.all_reduce_mean_scalar(torch.tensor(float(rank)))
    assert abs(s.item() - (world - 1) / 2) < 1e-6
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_gloo_world2_bucket_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"


def test_shard_range_partitions():
    from pde_policylearning_b200.parallel import shard_range
    for n in (1, 7, 64, 1024):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                seen += list(shard_range(n, r, w))
            assert seen == list(range(n))


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pde_policylearning_b200 import parallel
    parallel.init_from_env("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(), torch.nn.Linear(4, 2))
    extra = torch.nn.Parameter(torch.randn(3))                      # never reached by the loss: its bucket is zero-filled
    cw = torch.nn.Parameter(torch.randn(2, 2, dtype=torch.cfloat))
    params = list(net.parameters()) + [cw, extra]
    bucket = parallel.GradBucket(params)
    sync = parallel.OverlappedGradSync(bucket, bucket_bytes=64)      # tiny buckets: several groups
    assert len(sync.ranges) >= 3 and sync.ranges[0][1] == len(params)
    covered = sorted(i for lo, hi in sync.ranges for i in range(lo, hi))
    assert covered == list(range(len(params)))
    torch.manual_seed(10 + rank)
    for it in range(2):                                               # two steps: the hook state resets
        x = torch.randn(7, 6)
        for p in params:
            p.grad = None
        loss = net(x).square().sum() + (cw * (rank + 1)).abs().sum()
        loss.backward()
        assert sync.launched_during_backward > 0                      # groups were reduced before finish()
        local = [None if p.grad is None else p.grad.detach().clone() for p in params]
        # what the hooks have NOT touched yet must still be summed by finish()
        flat = sync.finish().clone()
        # reference: plain all-reduce of the local gradients
        for p, g in zip(params, local):
            g = torch.zeros_like(p) if g is None else g
            gr = torch.view_as_real(g).contiguous() if g.is_complex() else g.clone()
            dist.all_reduce(gr)
            got = torch.view_as_real(p.grad) if p.is_complex() else p.grad
            assert torch.allclose(got, gr, atol=1e-6), (it, tuple(p.shape))
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in params)
    if rank == 0:
        out.put("ok")
    dist.destroy_process_group()


def test_gloo_world2_overlapped_bucketed_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == "ok"
