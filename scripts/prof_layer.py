"""Runs the kernels of one FNO2d Fourier layer (BASELINE config 2 shape) a few times -- target for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_policylearning_b200 import ops

import os as _os
B, C, N, M = int(_os.environ.get("PROF_B", "64")), 32, 128, 12
dev = torch.device("cuda", 0)
plan = ops.get_plan(ops.SpecGeom(nin=(N, N), half=(M // 2, M // 2), norm="forward"), dev)
x = torch.randn(B, C, N, N, device=dev)
yh = torch.randn(B, C, *plan.kept, dtype=torch.complex64, device=dev)
w = torch.randn(C, C, device=dev)
bias = torch.randn(C, device=dev)
z = torch.randn_like(x)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for _ in range(iters):
    if what in ("all", "fwd"):
        xh = ops.dft_forward(plan, 0, x)
    if what in ("all", "inv"):
        y = ops.dft_inverse(plan, 0, yh, ops.make_epilogue(bias=bias, pw_w=w, pw_x=x))
    if what in ("all", "invgelu"):
        y = ops.dft_inverse(plan, 0, yh, ops.make_epilogue(bias=bias, pw_w=w, pw_x=x, preact=z, act="gelu"))
    if what in ("all", "inv3"):
        y = ops.dft_inverse(plan, 1, yh, ops.make_epilogue(pw_w=w, pw_x=x, pw_transposed=True, dact_z=z, dact="gelu"))
    if what in ("all", "wgrad"):
        ops.pw_wgrad(z, x, need_bias=False)
if what in ("all", "mlp"):
    import pde_policylearning_b200 as P
    w1 = (torch.randn(256, C, device=dev) * 0.2).requires_grad_(True)
    b1 = torch.randn(256, device=dev, requires_grad=True)
    w2 = (torch.randn(1, 256, device=dev) * 0.2).requires_grad_(True)
    b2 = torch.randn(1, device=dev, requires_grad=True)
    xr = x.clone().requires_grad_(True)
    for _ in range(iters):
        out = P.mlp_head(xr, w1, b1, w2, b2, "gelu")
        out.backward(torch.ones_like(out))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
if what == "time":
    import pde_policylearning_b200 as P
    w1 = (torch.randn(256, C, device=dev) * 0.2).requires_grad_(True)
    b1 = torch.randn(256, device=dev, requires_grad=True)
    w2 = (torch.randn(1, 256, device=dev) * 0.2).requires_grad_(True)
    b2 = torch.randn(1, device=dev, requires_grad=True)
    xr = x.clone().requires_grad_(True)
    gout = torch.ones(B, 1, N, N, device=dev)
    w1m, w2v = w1.detach(), w2.detach().reshape(-1)
    for name, fn in (("mlp_fwd", lambda: ops.mlp_head_fwd(x, w1m, b1.detach(), w2v, b2.detach(), "gelu")),
                     ("mlp_bwd", lambda: ops.mlp_head_bwd(x, w1m, b1.detach(), w2v, gout, "gelu", want_gz=True)),
                     ("mlp_bwd_nogz", lambda: ops.mlp_head_bwd(x, w1m, b1.detach(), w2v, gout, "gelu", want_gz=False)),
("fwd", lambda: ops.dft_forward(plan, 0, x)),
                     ("inv", lambda: ops.dft_inverse(plan, 0, yh, ops.make_epilogue(bias=bias, pw_w=w, pw_x=x))),
                     ("invgelu", lambda: ops.dft_inverse(plan, 0, yh, ops.make_epilogue(bias=bias, pw_w=w, pw_x=x, preact=z, act="gelu"))),
                     ("inv3", lambda: ops.dft_inverse(plan, 1, yh, ops.make_epilogue(pw_w=w, pw_x=x, pw_transposed=True, dact_z=z, dact="gelu"))),
                     ("wgrad", lambda: ops.pw_wgrad(z, x, need_bias=False))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(20):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        print(f"{name}: {ev[0].elapsed_time(ev[1]) / 20 * 1e3:.1f} us")
