#!/bin/bash
# gpurun call e: tests touched by the PINO kernel changes + full-size parity, PINO breakdown, cfg4 / cfg2 numbers, cfg2 launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for k in 3d_layer stages golden_pino golden_neuralop full_size fused_head cfg1; do
  timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  echo "[$k] rc=$? $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:|full size:" gpurun_out/pt_$k.log | head -8
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02e_pino_launches.csv python scripts/pino_step.py 4 1 > gpurun_out/r02e_pino.log 2>&1
python scripts/agg_launches.py gpurun_out/r02e_pino_launches.csv > gpurun_out/r02e_pino_breakdown.txt
head -16 gpurun_out/r02e_pino_breakdown.txt; tail -1 gpurun_out/r02e_pino.log
B2NO_SKIP_TF32=1 timeout 300 python bench.py --only cfg4 > gpurun_out/r02e_cfg4.json 2> gpurun_out/r02e_cfg4.err; cut -c1-500 gpurun_out/r02e_cfg4.json; tail -3 gpurun_out/r02e_cfg4.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 1 --warmup 3 --no-other --quick > gpurun_out/r02e_ncu_bench.log 2>&1
python scripts/step_breakdown.py gpurun_out/r02e_launches.csv > gpurun_out/r02e_step_breakdown.txt; cat gpurun_out/r02e_step_breakdown.txt | head -30
