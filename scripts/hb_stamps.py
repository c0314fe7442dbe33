"""Hand-over timeline of the one-kernel head backward (CTA 0, first steps).  Needs a library built with -DB2NO_HB_STAMPS:
   B2NO_LIB=/path/libb2no_stamps.so python scripts/hb_stamps.py"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_policylearning_b200 import ops, _lib

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, ci, hid, grid = 64, 32, 256, (128, 128)
x = torch.randn(B, ci, *grid, device=dev)
w1 = torch.randn(hid, ci, device=dev) * 0.2
b1 = torch.randn(hid, device=dev) * 0.1
w2 = torch.randn(hid, device=dev) * 0.2
g = torch.randn(B, 1, *grid, device=dev) * 1e-3
L = _lib.lib()
buf = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.mlp_head_bwd_fused(x, w1, b1, w2, g, "gelu")
torch.cuda.synchronize()
L.b2no_debug_head_bwd_stamps.argtypes = [ctypes.c_void_p]
assert L.b2no_debug_head_bwd_stamps(buf.data_ptr()) == 0
ops.mlp_head_bwd_fused(x, w1, b1, w2, g, "gelu")
torch.cuda.synchronize()
s = buf.cpu().view(3, 64, 8)
t0 = int(s[s > 0].min())
names = [["d1_empty ok", "G1 issued", "f_full ok", "G3 issued", "d2_empty ok", "G2 issued"],
         ["d1_full ok", "D1 read", "math done", "f_empty ok", "F written", "step end", "[gx: d2_full ok", "D2 read]"],
         ["raw_full ok", "xt_empty ok", "XT written", "xk_empty ok", "XK written"]]
for role, title in enumerate(["MMA issuer", "epilogue warp 6", "converter warp 2"]):
    print("==", title, "(cycles since first stamp)")
    for i in range(14 if role < 2 else 5):
        row = s[role, i]
        print("  %2d " % i + "  ".join("%s %7d" % (names[role][e], int(row[e]) - t0) for e in range(len(names[role])) if row[e] > 0))
