"""Timing of the head backward at the BASELINE config 2 shape: one-kernel path vs round 1's two-kernel path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_policylearning_b200 import ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, ci, hid, grid = 64, 32, 256, (128, 128)
xs = [torch.randn(B, ci, *grid, device=dev) for _ in range(3)]
w1 = torch.randn(hid, ci, device=dev) * 0.2
b1 = torch.randn(hid, device=dev) * 0.1
w2 = torch.randn(hid, device=dev) * 0.2
g = torch.randn(B, 1, *grid, device=dev) * 1e-3


def timeit(fn, n=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def fused(i):
    ops.mlp_head_bwd_fused(xs[i % 3], w1, b1, w2, g, "gelu")


def two(i):
    gx, gz, dw2 = ops.mlp_head_bwd(xs[i % 3], w1, b1, w2, g, "gelu", want_gz=True)
    ops.pw_wgrad(gz, xs[i % 3], need_bias=True)


b2 = torch.randn(1, device=dev)
print("head forward (k_mlp_tc<fwd, direct>): %.1f us" % timeit(lambda i: ops.mlp_head_fwd(xs[i % 3], w1, b1, w2, b2, "gelu")))
print("one-kernel head backward: %.1f us" % timeit(fused))
print("round-1 path (k_mlp_tc<bwd> + k_wgrad_tc): %.1f us" % timeit(two))
for mode in ("tf32",):
    ops.set_precision(mode) if hasattr(ops, "set_precision") else None
    print("one-kernel, single-pass TF32: %.1f us" % timeit(fused))
