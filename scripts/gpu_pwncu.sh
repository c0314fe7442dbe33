#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_pw_tc -s 2 -c 1 -f -o gpurun_out/pw1_ncu python scripts/prof_layer.py inv 4 > gpurun_out/pw1_ncu.log 2>&1
python scripts/ncu_top.py gpurun_out/pw1_ncu.ncu-rep 14 > gpurun_out/pw1_ncu.txt 2>&1; head -24 gpurun_out/pw1_ncu.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_pw_tc -s 2 -c 1 -f -o gpurun_out/pw3_ncu python scripts/prof_layer.py inv3 4 > gpurun_out/pw3_ncu.log 2>&1
python scripts/ncu_top.py gpurun_out/pw3_ncu.ncu-rep 14 > gpurun_out/pw3_ncu.txt 2>&1; head -8 gpurun_out/pw3_ncu.txt
