#!/bin/bash
mkdir -p gpurun_out
for k in 20 100 20; do timeout 300 python bench.py --quick --no-other --steps $k --warmup 5; done 2>&1 | tee gpurun_out/r2n_quick.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 1 --warmup 3 --no-other --quick > gpurun_out/r2n_ncu_bench.log 2>&1
python scripts/step_breakdown.py gpurun_out/r2n_launches.csv > gpurun_out/r2n_step_breakdown.txt; head -16 gpurun_out/r2n_step_breakdown.txt
