"""Locate the accuracy gap of the tiny FNO2dObserver smoke case: tensor-core vs CUDA-core path vs float64, layer by layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_policylearning_b200 as P
from pde_policylearning_b200 import ops
from oracle import closed_form as cf, restated as rs

dev = torch.device("cuda", 0)
torch.manual_seed(0)
conv = P.SpectralConv(32, 32, (16, 16), n_layers=1, factorization=None, implementation="factorized", fft_norm="forward")
_x = torch.randn(2, 32, 64, 64)
_gy = torch.randn(2, 32, 64, 64)
obs = P.FNO2dObserver(6, 6, 8).to(dev)
p = torch.randn(2, 16, 16, 1, device=dev)
sd = {k: (v.detach().cpu().to(torch.complex128) if v.is_complex() else v.detach().cpu().double()) for k, v in obs.state_dict().items()}
ref = rs.fno2d_observer_forward(sd, p.cpu().double(), 6)
for mode in (True, False):
    ops.set_tensor_core_mode(mode)
    out = obs(p)
    print("tc" if mode else "cc", "final", cf.rel_l2(out, ref))
    # layer by layer
    m = obs.fno2d
    grid = obs.get_grid(p.shape, dev)
    x = torch.cat((p, grid), dim=-1).permute(0, 3, 1, 2).contiguous()
    sub = {k[len("fno2d."):]: v for k, v in sd.items() if k.startswith("fno2d.")}
    x64 = x.cpu().double()
    h = m.lifting(x)
    h64 = torch.nn.functional.conv2d(x64, sub["lifting.fc.weight"], sub["lifting.fc.bias"])
    print("  lifting", cf.rel_l2(h, h64))
    for i in range(4):
        h = m.fno_blocks(h, i)
        corners = [sub[f"fno_blocks.convs.weight.{2*i+c}.tensor"] for c in range(2)]
        y = rs.neuralop_spectral_conv(h64, corners, sub["fno_blocks.convs.bias"][i].flatten(), (6, 6), "forward")
        y = y + torch.nn.functional.conv2d(h64, sub[f"fno_blocks.fno_skips.{i}.weight"])
        h64 = torch.nn.functional.gelu(y) if i < 4 - i else y
        print("  layer", i, cf.rel_l2(h, h64), "norm", h64.norm().item())
    o = m.projection(h)
    t = torch.nn.functional.gelu(torch.nn.functional.conv2d(h64, sub["projection.fc1.weight"], sub["projection.fc1.bias"]))
    o64 = torch.nn.functional.conv2d(t, sub["projection.fc2.weight"], sub["projection.fc2.bias"])
    print("  projection", cf.rel_l2(o, o64), "out norm", o64.norm().item(), "hidden norm", t.norm().item())
    # projection alone on exact input
    o2 = m.projection(h64.float().to(dev))
    print("  projection(exact input)", cf.rel_l2(o2, o64))
ops.set_tensor_core_mode(True)
