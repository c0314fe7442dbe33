#!/bin/bash
# full bench line (headline + other_configs) at N GPUs, as the driver launches it
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --master-port 29641 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/n${N}_bench.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu_baseline' in d)
for o in d.get('other_configs',[]): print(o.get('config'), o.get('value'), o.get('unit'), o.get('ms_per_step'), o.get('error'))
PY
tail -2 gpurun_out/n${N}_bench.err | cut -c1-200
