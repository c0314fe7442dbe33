#!/bin/bash
# gpurun call i: mode-major spectrum layout (tests, RNO launch list, cfg3 numbers), PINO stack path, full test suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for k in mode_major tensor_core_mixing regrouped golden_rno full_size golden_pinobserver fused_head; do
  timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  echo "[$k] rc=$? $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:|full size:" gpurun_out/pt_$k.log | cut -c1-300 | head -8
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02i_rno_launches.csv python scripts/rno_step.py 256 4 1 > gpurun_out/r02i_rno.log 2>&1
python scripts/agg_launches.py gpurun_out/r02i_rno_launches.csv > gpurun_out/r02i_rno_breakdown.txt
head -12 gpurun_out/r02i_rno_breakdown.txt
B2NO_CFG3_T=10 timeout 300 python bench.py --only cfg3 > gpurun_out/r02i_cfg3_T10.json 2> gpurun_out/r02i_cfg3_T10.err; python -c "
import json; d=json.load(open('gpurun_out/r02i_cfg3_T10.json')); print('cfg3 T=10', d['value'], d['ms_per_step'], d.get('tf32_mode'))"; tail -2 gpurun_out/r02i_cfg3_T10.err
B2NO_SKIP_TF32=1 timeout 300 python bench.py --only cfg4 > gpurun_out/r02i_cfg4.json 2> gpurun_out/r02i_cfg4.err; python -c "
import json; d=json.load(open('gpurun_out/r02i_cfg4.json')); print('cfg4', d['value'], d['ms_per_step'])"; tail -2 gpurun_out/r02i_cfg4.err
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --tb=line > gpurun_out/pt_all.log 2>&1; echo "[all] rc=$? $(tail -1 gpurun_out/pt_all.log)"; grep -E "^FAILED|^ERROR" gpurun_out/pt_all.log | head
