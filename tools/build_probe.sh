#!/bin/bash
# builds the stand-alone tcgen05 probes (sm_100a) next to their sources
set -e
cd "$(dirname "$0")"
for f in tc_probe; do
  [ -f $f.cu ] && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o $f.bin $f.cu
done
