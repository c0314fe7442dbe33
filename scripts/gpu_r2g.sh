#!/bin/bash
# gpurun call g: RNO launch list after the mixing-kernel changes, sanitizer passes, smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02g_rno_launches.csv python scripts/rno_step.py 256 4 1 > gpurun_out/r02g_rno.log 2>&1
python scripts/agg_launches.py gpurun_out/r02g_rno_launches.csv > gpurun_out/r02g_rno_breakdown.txt
head -14 gpurun_out/r02g_rno_breakdown.txt
B2NO_SKIP_TF32=1 B2NO_CFG3_T=10 timeout 300 python bench.py --only cfg3 > gpurun_out/r02g_cfg3_T10.json 2> gpurun_out/r02g_cfg3_T10.err; cut -c1-330 gpurun_out/r02g_cfg3_T10.json
bash scripts/gpu_sanitize.sh
