"""The fused PINO residual loss (csrc/pino_loss.cu) at the cfg4 shape, forward + backward -- ncu / timing target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_policylearning_b200 as P
dev = torch.device("cuda", 0)
B, N, T = 4, 64, 65
torch.manual_seed(0)
w = torch.randn(B, N, N, T, device=dev, requires_grad=True)
u0 = torch.randn(B, N, N, device=dev)
forcing = P.get_forcing(N, device=dev)
nu = 1.0 / torch.tensor([100.0, 200.0, 300.0, 400.0], device=dev)


def step():
    lic, lf = P.channelflow_pino_loss(w, u0, forcing, nu, 0.5)
    (lic + lf).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print(f"fused PINO residual loss fwd+bwd, B={B} {N}x{N}x{T}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call")
