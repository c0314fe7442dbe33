#!/bin/bash
# skeleton ablations (timing only)
mkdir -p gpurun_out
for m in 15 31 47 79 127; do echo "B2NO_HB_SKIP=$m"; B2NO_HB_SKIP=$m timeout 120 python scripts/hb_time.py 2>&1 | head -1; done | tee gpurun_out/hb_ablate2.log
