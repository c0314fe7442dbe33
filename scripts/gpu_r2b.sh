#!/bin/bash
# one gpurun call: new-kernel tests, the rest of the suite, cfg3 at T=4, PINO launch breakdown
bash scripts/gpu_r2.sh all
B2NO_CFG3_T=4 timeout 300 python bench.py --only cfg3 > gpurun_out/r02b_cfg3_T4.json 2> gpurun_out/r02b_cfg3_T4.err
cat gpurun_out/r02b_cfg3_T4.json; tail -5 gpurun_out/r02b_cfg3_T4.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02a_pino_launches.csv python scripts/pino_step.py 4 1 > gpurun_out/r02a_pino.log 2>&1
python scripts/agg_launches.py gpurun_out/r02a_pino_launches.csv > gpurun_out/r02a_pino_breakdown.txt
head -24 gpurun_out/r02a_pino_breakdown.txt; tail -2 gpurun_out/r02a_pino.log
