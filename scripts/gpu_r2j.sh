#!/bin/bash
# gpurun call j: fused 3-D inverse (tests, PINO breakdown, cfg4), then the whole suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for k in 3d_layer golden_pino cfg4_pino_full_size fused_head mirrors tensor_core_tile_kernel; do
  timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  echo "[$k] rc=$? $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:|full size:" gpurun_out/pt_$k.log | cut -c1-300 | head -8
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02j_pino_launches.csv python scripts/pino_step.py 4 1 > gpurun_out/r02j_pino.log 2>&1
python scripts/agg_launches.py gpurun_out/r02j_pino_launches.csv > gpurun_out/r02j_pino_breakdown.txt
head -16 gpurun_out/r02j_pino_breakdown.txt; tail -1 gpurun_out/r02j_pino.log
timeout 300 python bench.py --only cfg4 > gpurun_out/r02j_cfg4.json 2> gpurun_out/r02j_cfg4.err; python -c "
import json; d=json.load(open('gpurun_out/r02j_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'], d.get('tf32_mode'))"; tail -2 gpurun_out/r02j_cfg4.err
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --tb=line > gpurun_out/pt_all.log 2>&1; echo "[all] rc=$? $(tail -1 gpurun_out/pt_all.log)"; grep -E "^FAILED|^ERROR" gpurun_out/pt_all.log | head
