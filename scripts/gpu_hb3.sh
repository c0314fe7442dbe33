#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_head_bwd -s 3 -c 1 -o gpurun_out/hb_ncu2 -f python scripts/hb_time.py > gpurun_out/hb_ncu2.log 2>&1
tail -3 gpurun_out/hb_ncu2.log
