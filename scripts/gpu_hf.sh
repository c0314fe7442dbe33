#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 60 -k "mlp_head or head_backward or observer or fno2d or pino or rno" 2>&1 | tail -5 | tee gpurun_out/hf_test.log
timeout 100 python scripts/hb_time.py 2>&1 | tee gpurun_out/hf_time.log
