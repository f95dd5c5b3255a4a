#!/bin/bash
mkdir -p gpurun_out
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_k1d_f_tests.log 2>&1; tail -3 gpurun_out/r02_k1d_f_tests.log
for th in uniform mode; do
    python bench.py $Q --theta $th > gpurun_out/r02_k1d_f_$th.json 2>gpurun_out/r02_k1d_f.err
    python -c "
import json
d=json.load(open('gpurun_out/r02_k1d_f_$th.json')); print('$th', round(d['ms_per_step'],3), round(d['value']), d['clocks']['sm_mhz'])"
done
python scripts/numerics_table.py > gpurun_out/r02_numerics_k1d.txt 2>&1; tail -3 gpurun_out/r02_numerics_k1d.txt
