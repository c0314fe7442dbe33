#!/bin/bash
# parity + timing of the head kernels
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -s --timeout 40 -k "head_backward_one_kernel or tensor_core_mlp_head" 2>&1 | tail -9 | tee gpurun_out/hb_test.log
timeout 100 python scripts/hb_time.py 2>&1 | tee gpurun_out/hb_time.log
for m in 15 31; do echo "B2NO_HB_SKIP=$m"; B2NO_HB_SKIP=$m timeout 120 python scripts/hb_time.py 2>&1 | sed -n 2p; done | tee gpurun_out/hb_ablate.log
