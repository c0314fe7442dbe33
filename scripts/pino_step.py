"""One eager PINObserver2d training step (fwd + bwd, mean-square loss) at the cfg4 shape -- target for an ncu launch list.
usage: pino_step.py [B iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_policylearning_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = P.PINObserver2d(modes1=[8] * 4, modes2=[8] * 4, modes3=[8] * 4, fc_dim=128, layers=[64] * 5, act="gelu",
                    pad_ratio=0.0625).to(dev)
a = torch.randn(B, 64, 64, 65, 4, device=dev)
re = torch.rand(B, device=dev) * 400 + 100
for _ in range(iters):
    for p in m.parameters():
        p.grad = None
    out = m(a, re)
    out.square().mean().backward()
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
