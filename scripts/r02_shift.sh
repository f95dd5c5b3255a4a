#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or golden" 2>&1 | tail -2
python scripts/numerics_table.py 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-nuts --no-other-workloads --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', round(d['ms_per_step'],4))"
