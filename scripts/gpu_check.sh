#!/bin/bash
# Runs on the GPU box through gpurun: parity tests, smoke, bench, ncu launch list.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== stages" > gpurun_out/pytest.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 90 --tb=short -k "stages or epilogue or mlp_head" >> gpurun_out/pytest.log 2>&1
echo "== modules" >> gpurun_out/pytest.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 90 --tb=short -k "not (stages or epilogue or mlp_head)" >> gpurun_out/pytest.log 2>&1
timeout -s KILL 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
if [ "${1:-}" = "ncu" ]; then
  timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
fi
tail -5 gpurun_out/pytest.log
cat gpurun_out/smoke.log | tail -3
cat gpurun_out/bench.json | head -c 3000
