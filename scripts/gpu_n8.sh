#!/bin/bash
# N-GPU quick A/B of the gradient all-reduce placement (cfg2): plain (one all-reduce after backward) vs overlapped buckets
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for ov in 0 1; do
  B2NO_OVERLAP_AR=$ov timeout -s KILL 300 $TR --master-port 2962$ov bench.py --gpus $N --quick --no-other --steps 50 --warmup 5 > gpurun_out/n${N}_cfg2_ov$ov.json 2> gpurun_out/n${N}_cfg2_ov$ov.err
  echo "N=$N overlap=$ov rc=$? $(cut -c1-140 gpurun_out/n${N}_cfg2_ov$ov.json)"
done
