#!/bin/bash
# compute-sanitizer passes over the kernels of the second half of round 2: the one-kernel head backward (tc_head_bwd.cu), the
# two-group forward head and tile kernel (tc_mlp.cu, tc_pointwise.cu), the warp-per-plane forward DFT (spectral.cu)
mkdir -p gpurun_out
run() {  # tool name pytest-k
  timeout -s KILL 900 compute-sanitizer --tool $1 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 800 -k "$3" > gpurun_out/san2_$2.log 2>&1
  echo "[$1 $2] rc=$? $(grep -E 'passed|failed' gpurun_out/san2_$2.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san2_$2.log | tail -1)"
}
run memcheck head "head_backward_one_kernel or tensor_core_mlp_head"
run memcheck tile "tensor_core_tile_kernel or dft_forward_runs or stages_vs_closed_form"
run memcheck models "golden_fno2d or golden_rno or golden_pinobserver"
run racecheck head_race "head_backward_one_kernel"
run racecheck plane_race "dft_forward_runs or mode_major"
run synccheck head_sync "head_backward_one_kernel or tensor_core_mlp_head"
for f in head tile models head_race plane_race head_sync; do echo "== san2_$f"; grep -E "COMPUTE-SANITIZER|passed|failed|SUMMARY" gpurun_out/san2_$f.log | head -6; done > gpurun_out/san2_summary.txt
