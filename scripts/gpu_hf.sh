#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 8 9; do echo "B2NO_MLP_SKIP=$m"; B2NO_MLP_SKIP=$m timeout 100 python scripts/hb_time.py 2>&1 | head -1; done | tee gpurun_out/hf_ablate.log
