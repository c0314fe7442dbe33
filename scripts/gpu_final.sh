#!/bin/bash
# final-state record: smoke, full bench line, reference arm, ncu launch list of the bench step, RNO / PINO launch lists
mkdir -p gpurun_out
tag=${1:-r02k}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
for p in d['roofline']['all']:
    print('   %-70s %7.1f us frac %.3f share %.3f' % (p['kernel'][:70], p['seconds']*1e6, p['frac'], p['share_of_step']))
for o in d.get('other_configs',[]):
    print(o.get('config'), o.get('value'), o.get('unit'), o.get('ms_per_step'), 'tf32', o.get('tf32_mode',{}).get('value'), o.get('error'), o.get('wall_s'))
PY
tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-other --quick > gpurun_out/${tag}_ncu_bench.log 2>&1
python scripts/step_breakdown.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_step_breakdown.txt; head -14 gpurun_out/${tag}_step_breakdown.txt
