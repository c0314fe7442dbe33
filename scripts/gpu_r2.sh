#!/bin/bash
# round-2 GPU check: every new-kernel test in its OWN process (a sticky CUDA error or a hang in one must not hide the
# others), then optionally the rest of the suite.  Logs -> gpurun_out/pytest.log, one summary line per group on stdout.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
: > gpurun_out/pytest.log
for k in gate_epilogue 3d_layer tensor_core_mixing regrouped runs_on_tensor_cores golden_rno pino_residual golden_pinobserver mirrors per_sample_bias tf32_mode; do
  echo "== $k" >> gpurun_out/pytest.log
  timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  rc=$?
  cat gpurun_out/pt_$k.log >> gpurun_out/pytest.log
  echo "[$k] rc=$rc $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:" gpurun_out/pt_$k.log | head -6
done
if [ "${1:-}" = "all" ]; then
  echo "== rest" >> gpurun_out/pytest.log
  timeout -s KILL 700 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 150 --tb=short -k "not (gate_epilogue or 3d_layer or tensor_core_mixing or regrouped or runs_on_tensor_cores or golden_rno or pino_residual or golden_pinobserver or mirrors or per_sample_bias or tf32_mode)" > gpurun_out/pt_rest.log 2>&1
  cat gpurun_out/pt_rest.log >> gpurun_out/pytest.log
  echo "[rest] $(grep -E 'passed|failed|error' gpurun_out/pt_rest.log | tail -1)"
  grep -E "^FAILED|^ERROR" gpurun_out/pt_rest.log | head -10
fi
