#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: FNO fwd+bwd samples/s (+ SpectralConv roofline fraction).

Workload at every N (weak scaling): BASELINE config 2 -- FNO2dObserver(12,12,32) == FNO2d base_fno.yaml,
4 Fourier layers, 128x128 synthetic vorticity-like fields, batch 64 PER GPU, fp32.  One step = forward,
relative-L2 loss (run_pde_observers.py:138,188-193), backward, (N>1: one NCCL gradient all-reduce), Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM; `e2e` =
the same through the public module API with pinned HOST inputs (H2D + loss D2H inside the timed region).
`--impl reference` times the reference's CPU algorithm (oracle port; /root/reference does not travel to
the GPU box) on the host cores with all threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import warnings  # noqa: E402

import torch  # noqa: E402

warnings.filterwarnings("ignore")

METRIC = "fno2d_fwd_bwd_samples_per_s"
UNIT = "samples/s"
GRID = 128
MODES = 12
WIDTH = 32
BATCH = 64
LAYERS = 4


def config_dict(n_gpus):
    return {"workload": "FNO2d base_fno.yaml (FNO2dObserver modes 12, width 32, 4 layers, proj 256) on synthetic "
                        "128x128 fields, batch 64 per GPU, fp32, step = fwd + rel-L2 loss + bwd + Adam",
            "grid": GRID, "modes": MODES, "width": WIDTH, "layers": LAYERS, "batch_per_gpu": BATCH,
            "global_batch": BATCH * n_gpus, "parallelism": f"dp{n_gpus}",
            "l2_policy": "inputs larger than L2: every activation tensor is 134 MB (> 126 MB L2), ~5 GB touched per step"}


def synthetic_fields(batch, grid, seed, device="cpu"):
    """Smooth Gaussian random fields (spectrum ~ (k^2 + 49)^-1.25), unit variance: NS-vorticity-like."""
    g = torch.Generator().manual_seed(seed)
    k = torch.fft.fftfreq(grid, 1.0 / grid)
    k2 = k[:, None] ** 2 + k[None, :] ** 2
    amp = (k2 + 49.0) ** (-1.25)
    noise = torch.randn(batch, grid, grid, 2, generator=g)
    f = torch.fft.ifft2(torch.view_as_complex(noise) * amp).real
    f = f / f.std(dim=(1, 2), keepdim=True)
    return f.unsqueeze(-1).float().to(device)


# ---------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index=0, period=0.1):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.index, self.period = index, period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_batch):
    """Times fwd + rel-L2 + bwd + Adam of FNO2dObserver on the CPU with every host thread.
    Uses the unmodified reference when /root/reference exists (kind 'reference'), else the oracle port."""
    from oracle import ref_loader, restated as rs
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    p = synthetic_fields(sample_batch, GRID, seed=1)
    tgt = synthetic_fields(sample_batch, GRID, seed=2).permute(0, 3, 1, 2).contiguous()
    if ref_loader.available():
        kind = "reference"
        model = ref_loader.RefFNO2dObserver(MODES, MODES, WIDTH)
        params = list(model.parameters())
        fwd = lambda: model(p)
    else:
        kind = "port"
        import pde_policylearning_b200 as P
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in P.FNO2dObserver(MODES, MODES, WIDTH).state_dict().items()}
        params = list(sd.values())
        fwd = lambda: rs.fno2d_observer_forward(sd, p, MODES)
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = rs.lp_rel(fwd(), tgt, size_average=False)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_batch / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"batch {sample_batch} of the {BATCH}-sample step (same model, grid and step), "
                      f"{steps} timed steps after {warmup} warm-up, {dt * 1e3:.1f} ms/step",
            "ms_per_step": dt * 1e3}


# ---------------------------------------------------------------------------------------------
# per-kernel probes (CUDA events on the launching stream) for the roofline object
# ---------------------------------------------------------------------------------------------
def _time_op(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return sum(ts) / len(ts) * 1e-3  # seconds, mean


def kernel_probes(dev, hbm_peak_gbs):
    """Times the streaming kernels of one Fourier layer at the workload's shape, each alone, back to back
    over DIFFERENT buffers (working set 4 x 134 MB > L2)."""
    from pde_policylearning_b200 import ops
    B, C, N = BATCH, WIDTH, GRID
    geom = ops.SpecGeom(nin=(N, N), half=(MODES // 2, MODES // 2), norm="forward")
    plan = ops.get_plan(geom, dev)
    nbuf = 3
    xs = [torch.randn(B, C, N, N, device=dev) for _ in range(nbuf)]
    w = torch.randn(C, C, device=dev)
    bias = torch.randn(C, device=dev)
    spec = [torch.randn(B, C, *plan.kept, dtype=torch.complex64, device=dev) for _ in range(nbuf)]
    state = {"i": 0}
    L = ops._lib.lib()
    import ctypes as Cc
    work = plan.workspace(B, C)
    outs = [torch.empty(B, C, N, N, device=dev) for _ in range(nbuf)]
    A = torch.empty(B * C * N, plan.kept[1], dtype=torch.complex64, device=dev)
    st = Cc.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fwd_dft():
        i = state["i"] = (state["i"] + 1) % nbuf
        ops.check(L.b2no_dft_forward(plan.handle, 0, ops._ptr(xs[i]), ops._ptr(spec[i]), ops._ptr(work), B * C, st))

    def inv_fused():
        i = state["i"] = (state["i"] + 1) % nbuf
        epi = ops.make_epilogue(bias=bias, pw_w=w, pw_x=xs[i], act="gelu")
        ops.check(L.b2no_dft_inverse(plan.handle, 0, ops._ptr(spec[i]), ops._ptr(outs[i]), ops._ptr(work), B, C,
                                     N * N, Cc.byref(epi), st))

    def wgrad():
        i = state["i"] = (state["i"] + 1) % nbuf
        ops.pw_wgrad(xs[i], outs[(i + 1) % nbuf], need_bias=False)

    bytes_x = B * C * N * N * 4
    K = plan.modes
    probes = []
    # algorithmic bytes: forward DFT reads x once, writes the kept spectrum
    t = _time_op(fwd_dft)
    probes.append(dict(kernel="dft_forward (k_r2c_last + k_cmat)", seconds=t, bytes=bytes_x + B * C * K * 8,
                       launches_per_step=8))
    t = _time_op(inv_fused)
    probes.append(dict(kernel="dft_inverse fused (k_cmat + k_c2r_fused: irfft + bias + 1x1 skip + GELU)", seconds=t,
                       bytes=2 * bytes_x + B * C * K * 8, launches_per_step=8))
    t = _time_op(wgrad)
    probes.append(dict(kernel="pw_wgrad (1x1 skip weight gradient)", seconds=t, bytes=2 * bytes_x, launches_per_step=4))
    for p in probes:
        p["gbs"] = p["bytes"] / p["seconds"] / 1e9
        p["frac"] = p["gbs"] / hbm_peak_gbs
    return probes


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import pde_policylearning_b200 as P
    from pde_policylearning_b200 import ops, parallel
    import torch.distributed as dist

    rank, local_rank, world = parallel.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)                                    # identical weights on every rank
    model = P.FNO2dObserver(MODES, MODES, WIDTH).to(dev)
    loss_fn = P.LpLoss(size_average=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    bucket = parallel.GradBucket(model.parameters())
    bucket.attach()
    p_host = synthetic_fields(BATCH, GRID, seed=100 + rank).pin_memory()
    t_host = synthetic_fields(BATCH, GRID, seed=200 + rank).permute(0, 3, 1, 2).contiguous().pin_memory()
    p_dev, t_dev = p_host.to(dev), t_host.to(dev)

    def step(p, t):
        bucket.zero()
        loss = loss_fn(model(p, None), t)
        loss.backward()
        if world > 1:
            bucket.allreduce_mean()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(args.warmup):
        step(p_dev, t_dev)
    l0 = ops.LAUNCHES[0]
    with ClockSampler(local_rank) as clk:
        ms = timed(lambda: step(p_dev, t_dev), args.steps)
    launches = (ops.LAUNCHES[0] - l0) // args.steps
    clocks = clk.summary()
    ms_per_step = ms / args.steps
    value = BATCH * world / (ms_per_step * 1e-3)

    # ---- e2e: pinned host inputs -> H2D -> public API -> loss.item() each step ----
    def e2e_step():
        p = p_host.to(dev, non_blocking=True)
        t = t_host.to(dev, non_blocking=True)
        return step(p, t).item()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e = {"value": BATCH * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": p_host.numel() * 4 + t_host.numel() * 4, "d2h_bytes_per_step": 4}

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        probes = kernel_probes(dev, hbm)
        for p in probes:
            p["share_of_step"] = p["seconds"] * p["launches_per_step"] / (ms_per_step * 1e-3)
        top = max(probes, key=lambda p: p["share_of_step"])
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(top["kernel"])
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": round(top["gbs"], 1), "peak": hbm,
                    "unit": "GB/s", "frac": round(top["frac"], 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": top["bytes"], "avg_launch_us": round(top["seconds"] * 1e6, 2),
                    "all": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.items()} for p in probes]}
        cpu = None
        if world == 1 or True:
            cpu = cpu_reference_run(steps=3, warmup=1, sample_batch=4)
        out = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": config_dict(world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
               "roofline": roofline, "cpu_baseline": cpu}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    r = cpu_reference_run(steps=steps, warmup=warm, sample_batch=8)
    out = {"impl": "reference", "metric": METRIC, "value": round(r["value"], 3), "unit": UNIT, "n_gpus": args.gpus,
           "steps": steps, "warmup": warm, "ms_per_step": round(r["ms_per_step"], 2), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_dict(args.gpus),
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": round(r["value"], 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback (use --impl reference)")
        run_b200(args)


if __name__ == "__main__":
    main()
