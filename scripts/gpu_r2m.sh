#!/bin/bash
# round 2, second session: full GPU suite + head-backward timing + bench with the one-kernel head backward
mkdir -p gpurun_out
timeout 100 python scripts/hb_time.py 2>&1 | tee gpurun_out/r2m_hb_time.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -15 | tee gpurun_out/r2m_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
tail -c 600 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
for p in d['roofline']['all']: print('   %-75s %7.1f us frac %.3f share %.3f' % (p['kernel'][:75], p['seconds']*1e6, p['frac'], p['share_of_step']))
for o in d.get('other_configs', []): print(o.get('config',{}).get('workload','?')[:40], o.get('value'), o.get('unit'), o.get('ms_per_step'))
PY
