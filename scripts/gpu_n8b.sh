#!/bin/bash
# N-GPU quick check of the headline step (plain all-reduce)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 200 $TR --master-port 29631 bench.py --gpus $N --quick --no-other --steps 20 --warmup 5 > gpurun_out/n${N}_cfg2.json 2> gpurun_out/n${N}_cfg2.err
echo "N=$N rc=$? $(cut -c1-140 gpurun_out/n${N}_cfg2.json)"; tail -2 gpurun_out/n${N}_cfg2.err | cut -c1-200
