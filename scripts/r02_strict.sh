#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small.py tests/test_gpu_parity.py -m gpu -x -q -k "strict or near_the_mode" 2>&1 | tail -3
Q="--steps 10 --warmup 3 --no-nuts --no-other-workloads --no-cpu-baseline"
for e in 0 1; do
  BL_STRICT_ENGINE=$e python bench.py $Q --strict-math 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('strict math, BL_STRICT_ENGINE=$e: ms_per_step', round(d['ms_per_step'],3), 'value', round(d['value']), 'roofline', d.get('roofline',{}).get('frac'))"
done
python scripts/numerics_table.py 2>&1 | tail -12
