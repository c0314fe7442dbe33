#!/bin/bash
# round 2 (second session): parity + timing + hand-over timeline of the one-kernel head backward
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -s --timeout 40 -k "head_backward_one_kernel or tensor_core_mlp_head" 2>&1 | tail -12 | tee gpurun_out/hb_test.log
timeout 100 python scripts/hb_time.py 2>&1 | tee gpurun_out/hb_time.log
B2NO_LIB=$PWD/pde_policylearning_b200/libb2no_stamps.so timeout 100 python scripts/hb_stamps.py 2>&1 | tee gpurun_out/hb_stamps.log
for m in 7 8 15; do echo "B2NO_HB_SKIP=$m"; B2NO_HB_SKIP=$m timeout 120 python scripts/hb_time.py 2>&1 | head -1; done | tee gpurun_out/hb_ablate.log
