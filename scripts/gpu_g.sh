#!/bin/bash
mkdir -p gpurun_out
timeout 100 python scripts/hb_time.py 2>&1 | tee gpurun_out/g_time.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4
for k in 1 2; do timeout 300 python bench.py --quick --no-other --steps 50 --warmup 5; done
