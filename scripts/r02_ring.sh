#!/bin/bash
# K1d ring_mode A/B: parity at full size with the barrier-free ring, then ms per 1024-chain evaluation in both modes
Q="--steps 20 --warmup 5 --no-nuts --no-other-workloads --no-cpu-baseline"
BL_SIGNED_RING=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2 or golden" 2>&1 | tail -2
for m in 0 1 0 1; do
  for th in init mode; do
    BL_SIGNED_RING=$m python bench.py $Q --theta $th 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ring_mode=$m theta=$th ms_per_step', round(d['ms_per_step'],4))"
  done
done
