#!/bin/bash
mkdir -p gpurun_out
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_k1d_e_tests.log 2>&1; tail -2 gpurun_out/r02_k1d_e_tests.log
for v in 1 2; do
  for th in uniform mode; do
    BL_SIGNED_NS=$v python bench.py $Q --theta $th > gpurun_out/r02_k1d_e_v${v}_$th.json 2>gpurun_out/r02_k1d_e.err
    python -c "
import json
d=json.load(open('gpurun_out/r02_k1d_e_v${v}_$th.json')); print('NS=$v $th', round(d['ms_per_step'],3), round(d['value']), d['clocks']['sm_mhz'])"
  done
done
python scripts/numerics_table.py > gpurun_out/r02_numerics_k1d.txt 2>&1; tail -12 gpurun_out/r02_numerics_k1d.txt
