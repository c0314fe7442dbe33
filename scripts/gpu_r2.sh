#!/bin/bash
# round-2 GPU check: new-kernel tests first, then the rest, RNO/PINO timings, bench.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
tag=${1:-r02}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== new" > gpurun_out/pytest.log
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 --tb=short -k "mixing or gate_epilogue or regrouped or runs_on_tensor_cores or golden_rno" >> gpurun_out/pytest.log 2>&1
tail -15 gpurun_out/pytest.log
if [ "${2:-}" = "all" ]; then
  echo "== rest" >> gpurun_out/pytest.log
  timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 --tb=short -k "not (mixing or gate_epilogue or regrouped or runs_on_tensor_cores or golden_rno)" >> gpurun_out/pytest.log 2>&1
  tail -5 gpurun_out/pytest.log
fi
