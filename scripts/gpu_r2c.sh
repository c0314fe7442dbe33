#!/bin/bash
# gpurun call c: new tests, RNO launch breakdown, PINO breakdown with the split layer + fused head, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for k in 3d_layer fused_head golden_pinobserver golden_pino_conv; do
  timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  echo "[$k] rc=$? $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:" gpurun_out/pt_$k.log | head -6
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02c_rno_launches.csv python scripts/rno_step.py 256 4 1 > gpurun_out/r02c_rno.log 2>&1
python scripts/agg_launches.py gpurun_out/r02c_rno_launches.csv > gpurun_out/r02c_rno_breakdown.txt
head -32 gpurun_out/r02c_rno_breakdown.txt; tail -2 gpurun_out/r02c_rno.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02c_pino_launches.csv python scripts/pino_step.py 4 1 > gpurun_out/r02c_pino.log 2>&1
python scripts/agg_launches.py gpurun_out/r02c_pino_launches.csv > gpurun_out/r02c_pino_breakdown.txt
head -20 gpurun_out/r02c_pino_breakdown.txt; tail -2 gpurun_out/r02c_pino.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"; cat gpurun_out/r02c_bench.json | head -c 6000; tail -5 gpurun_out/r02c_bench.err
