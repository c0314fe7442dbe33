#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (small test cases): memcheck on the tensor-core mixing / gate / 3-D /
# PINO-loss tests, racecheck on the shared-memory kernels of the PINO residual loss and the last-dim stages.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() {  # tool name pytest-k
  timeout -s KILL 600 compute-sanitizer --tool $1 --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 -k "$3" > gpurun_out/san_$2.log 2>&1
  echo "[$1 $2] rc=$? $(grep -E 'passed|failed' gpurun_out/san_$2.log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$2.log | tail -1)"
}
run memcheck mix "tensor_core_mixing"
run memcheck gate "gate_epilogue or single_a_buffer"
run memcheck pino "pino_residual or 3d_layer"
run memcheck rno "golden_rno"
run racecheck pino_race "pino_residual"
run racecheck stages_race "3d_layer"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
