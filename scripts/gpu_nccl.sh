#!/bin/bash
# all-reduce tail of the cfg2 step at N GPUs under different NCCL protocol choices (2.4 MB message)
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for env in "" "NCCL_PROTO=LL" "NCCL_PROTO=LL128" "NCCL_PROTO=Simple" "NCCL_ALGO=Tree" "NCCL_MAX_NCHANNELS=4" "NCCL_MIN_NCHANNELS=16"; do
  i=$((i+1))
  out=$(env $env timeout -s KILL 200 $TR --master-port $((29650+i)) bench.py --gpus $N --quick --no-other --steps 50 --warmup 5 2>/dev/null | tail -1 | cut -c40-110)
  echo "[$env] $out"
done | tee gpurun_out/nccl_n$N.log
