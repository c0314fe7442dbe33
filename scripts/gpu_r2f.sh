#!/bin/bash
# gpurun call f: PINO kernel changes (tests + breakdown), cfg2 quick after the MMA-loop fix, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for k in 3d_layer stages golden_pino full_size tensor_core_tile_kernel golden_fno2d tf32_mode; do
  timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  echo "[$k] rc=$? $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:|full size:" gpurun_out/pt_$k.log | cut -c1-400 | head -8
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02f_pino_launches.csv python scripts/pino_step.py 4 1 > gpurun_out/r02f_pino.log 2>&1
python scripts/agg_launches.py gpurun_out/r02f_pino_launches.csv > gpurun_out/r02f_pino_breakdown.txt
head -16 gpurun_out/r02f_pino_breakdown.txt; tail -1 gpurun_out/r02f_pino.log
timeout 300 python bench.py --quick --no-other > gpurun_out/r02f_cfg2_quick.json 2> gpurun_out/r02f_cfg2_quick.err; cat gpurun_out/r02f_cfg2_quick.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
for o in d.get('other_configs',[]):
    print(o.get('config'), o.get('value'), o.get('unit'), o.get('ms_per_step'), o.get('tf32_mode',{}).get('value'), o.get('error'), o.get('wall_s'))
PY
tail -3 gpurun_out/r02f_bench.err
