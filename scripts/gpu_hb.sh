#!/bin/bash
# round 2 (second session): parity + timing + hand-over timeline + ncu of the one-kernel head backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -s -k "head_backward_one_kernel or tensor_core_mlp_head" 2>&1 | tail -40 | tee gpurun_out/hb_test.log
timeout 300 python scripts/hb_time.py 2>&1 | tee gpurun_out/hb_time.log
B2NO_LIB=$PWD/pde_policylearning_b200/libb2no_stamps.so timeout 300 python scripts/hb_stamps.py 2>&1 | tee gpurun_out/hb_stamps.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_head_bwd -s 3 -c 1 -o gpurun_out/hb_ncu -f python scripts/hb_time.py > gpurun_out/hb_ncu.log 2>&1
