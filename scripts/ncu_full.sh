#!/bin/bash
# ncu --set full captures of the hot kernels (one GPU).  usage: ncu_full.sh <tag> [what...]; reports -> gpurun_out/<tag>_<what>.ncu-rep
tag=${1:-r01}; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    mlp)   k="k_mlp_tc"; skip=2; cnt=2;;
    fwd)   k="k_r2c_last|k_cmat|k_fwd_tc"; skip=1; cnt=1;;
    inv)   k="k_pw_tc|k_inv_h"; skip=2; cnt=2;;
    invgelu) k="k_pw_tc"; skip=1; cnt=1;;
    inv3)  k="k_pw_tc"; skip=1; cnt=1;;
    wgrad) k="k_wgrad_tc"; skip=1; cnt=1;;
    invh)  k="k_inv_h"; skip=1; cnt=1; what2=inv;;
  esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c $cnt -f -o gpurun_out/${tag}_${what} \
      python scripts/prof_layer.py ${what2:-$what} 2 > gpurun_out/${tag}_${what}.log 2>&1
  echo "$what exit $?"
done
