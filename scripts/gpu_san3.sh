#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -k "head_backward_one_kernel" > gpurun_out/san2_head_race.log 2>&1
echo "rc=$? $(grep -E 'passed|failed' gpurun_out/san2_head_race.log | tail -1) | $(grep -E 'RACECHECK SUMMARY' gpurun_out/san2_head_race.log | tail -1)"
grep -n "Error: Race" -A3 gpurun_out/san2_head_race.log | head -20
for f in head tile models head_race plane_race head_sync; do echo "== san2_$f"; grep -E "COMPUTE-SANITIZER|passed|failed|SUMMARY" gpurun_out/san2_$f.log | head -6; done > gpurun_out/san2_summary.txt
timeout 100 python scripts/hb_time.py | head -3
