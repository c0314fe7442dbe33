#!/bin/bash
# round-2 `ncu --set full` captures (one GPU): one launch of each hot kernel of the RNO step (cfg3 layer shape), the PINO
# step (cfg4 shape) and the cfg2 head backward.  Reports -> gpurun_out/<tag>_<name>.ncu-rep, text summaries next to them.
tag=${1:-r02d}
mkdir -p gpurun_out
cap() {  # name regex skip script args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c 1 -f -o gpurun_out/${tag}_${name} "$@" > gpurun_out/${tag}_${name}.log 2>&1
  echo "$name exit $?"
  python scripts/ncu_top.py gpurun_out/${tag}_${name}.ncu-rep 14 > gpurun_out/${tag}_ncu_${name}.txt 2>&1
}
cap rno_mix   "k_mix_tc"   6 python scripts/rno_step.py 256 2 1
cap rno_dw    "k_dw_tc"    2 python scripts/rno_step.py 256 2 1
cap rno_pw0   "k_pw_tc<0"  6 python scripts/rno_step.py 256 2 1
cap rno_fwd   "k_fwd_tc"   6 python scripts/rno_step.py 256 2 1
cap rno_invh  "k_inv_h"    6 python scripts/rno_step.py 256 2 1
cap pino_c2r  "k_c2r_fused" 2 python scripts/pino_step.py 4 1
cap pino_r2c  "k_r2c_last" 2 python scripts/pino_step.py 4 1
cap pino_res  "k_pino_residual_fwd" 0 python scripts/pino_loss_step.py
cap pino_resb "k_pino_residual_bwd" 0 python scripts/pino_loss_step.py
cap mlp_bwd   "k_mlp_tc"   3 python scripts/prof_layer.py mlp 2
ls -la gpurun_out/${tag}_*.ncu-rep | awk '{print $5, $9}'
