#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rn or occu_rn or ragged or refbody or golden" > gpurun_out/r02_rn2_tests.log 2>&1; tail -15 gpurun_out/r02_rn2_tests.log
Q="--workload occu_rn_200k_x10_k50 --no-nuts --no-cpu-baseline --steps 5"
for v in "BL_RN_CHAIN_KERNEL=1" "BL_RN2_BT=256" "BL_RN2_BT=128"; do
  for th in uniform mode; do
    env $v python bench.py $Q --theta $th > gpurun_out/r02_rn2_tmp.json 2>gpurun_out/r02_rn2.err || tail -3 gpurun_out/r02_rn2.err
    python -c "
import json
d=json.load(open('gpurun_out/r02_rn2_tmp.json')); print('$v $th', round(d['ms_per_step'],3), round(d['value']), d['clocks']['sm_mhz'])"
  done
done
