#!/bin/bash
mkdir -p gpurun_out
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_k1d_c_tests.log 2>&1; tail -2 gpurun_out/r02_k1d_c_tests.log
for g in 1 2 4 8; do
  for v in 0 21; do
    BL_SIGNED_G=$g BL_SIGNED_NS=$v python bench.py $Q > gpurun_out/r02_k1d_g${g}_v$v.json 2>gpurun_out/r02_k1d_c.err
    python -c "
import json
d=json.load(open('gpurun_out/r02_k1d_g${g}_v$v.json')); print('G=$g variant $v', round(d['ms_per_step'],3), round(d['value']), d['clocks']['sm_mhz'])"
  done
done
BL_SIGNED_G=4 python bench.py $Q --theta mode | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('G=4 mode', round(d['ms_per_step'],3))"
