"""Timing of the OTHER BASELINE configs on the current kernels (not bench lines: round-2 starting points).
cfg3: RNO observer (configs/matlab_rno.yaml shape: width 34, modes 12, 32x32 planes), fwd + rel-L2 + bwd, eager.
cfg4: PINObserver2d (pino-observer-pretrain-1s.yaml shape: 4 layers x 64 ch, modes 8, 64x64x65 grid), fwd + bwd, eager.
usage: bench_other.py [rno_B rno_T pino_B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_policylearning_b200 as P

dev = torch.device("cuda", 0)
rno_B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rno_T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pino_B = int(sys.argv[3]) if len(sys.argv) > 3 else 1


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
try:
    m = P.RNO2dObserver(12, 12, 34, 0, layer_num=1).to(dev).eval()      # eval: the reference's dropout(0.3) off, as in the parity tests
    x = torch.randn(rno_B, rno_T, 32, 32, 1, device=dev)
    tgt = torch.randn(rno_B, 32, 32, 1, device=dev)

    def rno_step():
        for p in m.parameters():
            p.grad = None
        out = m(x)
        loss = P.rel_l2_loss(out.reshape(rno_B, -1), tgt.reshape(rno_B, -1), size_average=False)
        loss.backward()

    ms = timed(rno_step)
    print(f"cfg3 RNO2dObserver(12,12,34,L=1) B={rno_B} T={rno_T} 32x32 fwd+bwd eager: {ms:.1f} ms/step = "
          f"{rno_B / ms * 1e3:.1f} trajectories/s at T={rno_T} ({ms / rno_T:.2f} ms per recurrent step)")
    # the same step (plus fused Adam) captured once into a CUDA graph: what is left when the host launch loop is gone
    opt = P.FusedAdam(m.parameters(), lr=1e-3)
    lf = lambda o, t: P.rel_l2_loss(o.reshape(rno_B, -1), t.reshape(rno_B, -1), size_average=False)
    gstep = P.GraphedTrainStep(m, lf, opt, (x,), tgt, warmup=2)
    ms = timed(lambda: gstep((x,), tgt))
    print(f"cfg3 same, fwd + loss + bwd + Adam as ONE CUDA graph ({gstep.launches_per_step} library launches): {ms:.1f} ms/step = "
          f"{rno_B / ms * 1e3:.1f} trajectories/s at T={rno_T} ({ms / rno_T:.2f} ms per recurrent step)")
    gstep.close()
except Exception as e:  # noqa: BLE001
    print("cfg3 failed:", type(e).__name__, e)

try:
    m = P.PINObserver2d(modes1=[8] * 4, modes2=[8] * 4, modes3=[8] * 4, fc_dim=128, layers=[64] * 5, act="gelu",
                        pad_ratio=0.0625).to(dev)
    a = torch.randn(pino_B, 64, 64, 65, 4, device=dev)
    re = torch.rand(pino_B, device=dev) * 400 + 100

    def pino_step():
        for p in m.parameters():
            p.grad = None
        out = m(a, re)
        out.square().mean().backward()

    ms = timed(pino_step)
    print(f"cfg4 PINObserver2d(4 x 64 ch, modes 8) B={pino_B} 64x64x65 fwd+bwd eager (one forward, mean-square loss): "
          f"{ms:.1f} ms/step = {pino_B / ms * 1e3:.2f} samples/s")
except Exception as e:  # noqa: BLE001
    print("cfg4 failed:", type(e).__name__, e)
