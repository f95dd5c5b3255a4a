#!/bin/bash
# K1d pipeline / occupancy variants on config 2
mkdir -p gpurun_out
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_k1d_b_tests.log 2>&1; tail -2 gpurun_out/r02_k1d_b_tests.log
for v in 0 20 1 13 23; do
  for th in uniform mode; do
    BL_SIGNED_NS=$v python bench.py $Q --theta $th > gpurun_out/r02_k1d_v${v}_$th.json 2>gpurun_out/r02_k1d_b.err
    python -c "
import json
d=json.load(open('gpurun_out/r02_k1d_v${v}_$th.json')); print('variant $v $th', round(d['ms_per_step'],3), round(d['value']), d['clocks']['sm_mhz'])"
  done
done
