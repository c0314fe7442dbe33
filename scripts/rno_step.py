"""One eager RNO observer training step (fwd + rel-L2 + bwd) at the cfg3 layer shape -- target for an ncu launch list.
usage: rno_step.py [B T iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_policylearning_b200 as P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = P.RNO2dObserver(12, 12, 34, 0, layer_num=1).to(dev).eval()
x = torch.randn(B, T, 32, 32, 1, device=dev)
tgt = torch.randn(B, 32, 32, 1, device=dev)
for _ in range(iters):
    for p in m.parameters():
        p.grad = None
    loss = P.rel_l2_loss(m(x).reshape(B, -1), tgt.reshape(B, -1), size_average=False)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
