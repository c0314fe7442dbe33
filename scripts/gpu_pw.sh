#!/bin/bash
mkdir -p gpurun_out
for m in 0 3; do echo "B2NO_TC_DEBUG=$m"; B2NO_TC_DEBUG=$m timeout 120 python scripts/prof_layer.py time 2>&1 | grep -E "^(mlp_fwd|fwd|inv|invgelu|inv3|wgrad):"; done | tee gpurun_out/pw_ablate2.log
timeout 100 python scripts/hb_time.py 2>&1 | head -2
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -4
