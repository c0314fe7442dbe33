import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B2NO_FD_DEBUG"] = "16"
import torch
from pde_policylearning_b200 import ops, _lib
B = int(os.environ.get("PROF_B", "64"))
dev = torch.device("cuda", 0)
plan = ops.get_plan(ops.SpecGeom(nin=(128, 128), half=(6, 6), norm="forward"), dev)
x = torch.randn(B, 32, 128, 128, device=dev)
for _ in range(3):
    ops.dft_forward(plan, 0, x)
torch.cuda.synchronize()
ts = (ctypes.c_ulonglong * 16)()
_lib.lib().b2no_debug_fd_ts(ts)
t0 = ts[0]
names = ["start", "setup done", "first TMA box landed", "first chunk in TMEM", "Mt loaded", "first d1_full", "first handover done",
         "first d2_full", "epilogue loop done", "exit"]
for n, t in zip(names, ts):
    print(f"{n:28s} {(t - t0) / 1000:8.2f} us")
