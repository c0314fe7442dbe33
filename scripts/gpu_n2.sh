#!/bin/bash
# 2-GPU call: full bench line at N=2, then the overlapped all-reduce against the plain one (cfg2 quick, cfg4)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02h_bench_n2.json'))
    print(d['value'], d['ms_per_step'], d['e2e']['value'])
    for o in d.get('other_configs',[]):
        print(o.get('config'), o.get('value'), o.get('unit'), o.get('ms_per_step'), o.get('tf32_mode',{}).get('value'), o.get('error'))
except Exception as e:
    print("no json", e)
PY
tail -3 gpurun_out/r02h_bench_n2.err
for ov in 0 1; do
  B2NO_OVERLAP_AR=$ov timeout -s KILL 300 $TR --master-port 2952$ov bench.py --gpus 2 --quick --no-other > gpurun_out/r02h_cfg2_ov$ov.json 2> gpurun_out/r02h_cfg2_ov$ov.err
  echo "cfg2 overlap=$ov rc=$? $(cat gpurun_out/r02h_cfg2_ov$ov.json | cut -c1-160)"; tail -2 gpurun_out/r02h_cfg2_ov$ov.err
  B2NO_SKIP_TF32=1 B2NO_OVERLAP_AR=$ov timeout -s KILL 300 $TR --master-port 2953$ov bench.py --gpus 2 --only cfg4 > gpurun_out/r02h_cfg4_ov$ov.json 2> gpurun_out/r02h_cfg4_ov$ov.err
  echo "cfg4 overlap=$ov rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r02h_cfg4_ov$ov.json'));print(d['value'],d['ms_per_step'])" 2>&1 | tail -1)"; tail -2 gpurun_out/r02h_cfg4_ov$ov.err
done
