#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 -k "stages or golden or tile_kernel or rno_layer" 2>&1 | tail -2
timeout 120 python scripts/prof_layer.py time 2>&1 | grep -E "^(inv|invgelu|inv3):"
for k in 1 2; do timeout 300 python bench.py --quick --no-other --steps 50 --warmup 5 | cut -c40-100; done
timeout 300 python bench.py --only cfg3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3', d.get('value'), d.get('ms_per_step'))"
