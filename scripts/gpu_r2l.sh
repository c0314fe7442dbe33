#!/bin/bash
# record of the head kernels (ncu --set full summaries, hand-over timeline, ablations) followed by the final-state record
# (scripts/gpu_final.sh: smoke, bench line, reference arm, ncu launch list + step breakdown)
mkdir -p gpurun_out
tag=${1:-r02m}
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_head_bwd -s 3 -c 1 -f -o gpurun_out/${tag}_head_bwd python scripts/hb_time.py > gpurun_out/${tag}_ncu_head_bwd.log 2>&1
python scripts/ncu_top.py gpurun_out/${tag}_head_bwd.ncu-rep 16 > gpurun_out/${tag}_ncu_head_bwd.txt 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_mlp_tc -s 3 -c 1 -f -o gpurun_out/${tag}_head_fwd python scripts/hb_time.py > gpurun_out/${tag}_ncu_head_fwd.log 2>&1
python scripts/ncu_top.py gpurun_out/${tag}_head_fwd.ncu-rep 16 > gpurun_out/${tag}_ncu_head_fwd.txt 2>&1
timeout 100 python scripts/hb_time.py 2>&1 | tee gpurun_out/${tag}_head_time.log
for m in 7 8 15 31 47; do echo "B2NO_HB_SKIP=$m"; B2NO_HB_SKIP=$m timeout 120 python scripts/hb_time.py 2>&1 | sed -n 2p; done | tee gpurun_out/${tag}_head_ablate.log
timeout 120 python scripts/prof_layer.py time 2>&1 | grep -E "^(mlp_fwd|fwd|inv|invgelu|inv3|wgrad):" | tee gpurun_out/${tag}_layer_time.log
B2NO_LIB=$PWD/pde_policylearning_b200/libb2no_stamps.so timeout 100 python scripts/hb_stamps.py > gpurun_out/${tag}_head_stamps.log 2>&1
bash scripts/gpu_final.sh $tag
