#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q --timeout 120 -k "rno or mode_major or dft or golden" 2>&1 | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/rno_launches.csv python scripts/rno_step.py 256 4 2 > gpurun_out/rno_ncu.log 2>&1
python scripts/agg_launches.py gpurun_out/rno_launches.csv 2>/dev/null | head -6
