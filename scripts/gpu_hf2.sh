#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_mlp_tc -s 3 -c 1 -f -o gpurun_out/hf_ncu python scripts/hb_time.py > gpurun_out/hf_ncu.log 2>&1
python scripts/ncu_top.py gpurun_out/hf_ncu.ncu-rep 30 > gpurun_out/hf_ncu.txt 2>&1; head -60 gpurun_out/hf_ncu.txt
