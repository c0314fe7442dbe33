"""Reads the %globaltimer stamps k_wgrad_tc records for CTA 0 (B2NO_WG_DEBUG=16): where a chunk's hand-over time goes.
Needs a library built with the stamps compiled in:
    python -c "import subprocess; from pde_policylearning_b200 import _lib; subprocess.run(_lib.nvcc_command('/tmp/libb2no_stamps.so', ['B2NO_WG_STAMPS']), check=True)"
    B2NO_LIB=/tmp/libb2no_stamps.so python scripts/wg_ts.py
"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B2NO_WG_DEBUG"] = "16"
import torch
from pde_policylearning_b200 import ops, _lib
B = int(os.environ.get("PROF_B", "64"))
dev = torch.device("cuda", 0)
g = torch.randn(B, 32, 128, 128, device=dev)
x = torch.randn(B, 32, 128, 128, device=dev)
for _ in range(3):
    ops.pw_wgrad(g, x, need_bias=False)
torch.cuda.synchronize()
ts = (ctypes.c_ulonglong * 128)()
L = _lib.lib()
L.b2no_debug_wg_ts.argtypes = [ctypes.c_void_p]
L.b2no_debug_wg_ts(ts)
t0 = min(t for t in ts if t)
names = ["slot free", "lo done", "conv issued", "st complete", "mma saw full", "mma issued", "stage full", "-"]
for n in range(12):
    print(f"chunk {n:2d}: " + "  ".join(f"{names[k]} {((ts[n * 8 + k] - t0) / 1000 if ts[n * 8 + k] else -1):7.2f}" for k in (6, 0, 1, 2, 3, 4, 5)))
