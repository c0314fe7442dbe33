#!/bin/bash
# gpurun call d: tests of the changed kernels, quick timings (fp32 / tf32 mode), ncu full captures
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for k in mode_major tensor_core_mixing regrouped golden_rno cfg3_rno_full_size gate_epilogue; do
  timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 --tb=short -s -k "$k" > gpurun_out/pt_$k.log 2>&1
  echo "[$k] rc=$? $(grep -E 'passed|failed|error' gpurun_out/pt_$k.log | tail -1)"
  grep -E "^E  |Error|error:" gpurun_out/pt_$k.log | head -6
done
timeout 120 python scripts/pino_loss_step.py
B2NO_CFG3_T=10 timeout 300 python bench.py --only cfg3 > gpurun_out/r02d_cfg3_T10.json 2> gpurun_out/r02d_cfg3_T10.err; cut -c1-700 gpurun_out/r02d_cfg3_T10.json; tail -3 gpurun_out/r02d_cfg3_T10.err
timeout 300 python bench.py --only cfg4 > gpurun_out/r02d_cfg4.json 2> gpurun_out/r02d_cfg4.err; cut -c1-700 gpurun_out/r02d_cfg4.json; tail -3 gpurun_out/r02d_cfg4.err
timeout 300 python bench.py --quick --no-other > gpurun_out/r02d_cfg2_quick.json 2> gpurun_out/r02d_cfg2_quick.err; cat gpurun_out/r02d_cfg2_quick.json
bash scripts/ncu_r2.sh r02d
