#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_k1d_d_tests.log 2>&1; tail -2 gpurun_out/r02_k1d_d_tests.log
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
python bench.py $Q > gpurun_out/r02_k1d_final.json 2>gpurun_out/r02_k1d_d.err; python -c "
import json
d=json.load(open('gpurun_out/r02_k1d_final.json')); print('final', round(d['ms_per_step'],3), round(d['value']), d['clocks'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:occu_signed_kernel -s 3 -c 1 -o gpurun_out/r02_k1d_v2 python bench.py --steps 1 --warmup 3 --no-nuts --no-other-workloads --no-cpu-baseline > gpurun_out/r02_k1d_ncu2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv --log-file gpurun_out/r02_launches_k1d.csv python bench.py --steps 5 --warmup 3 --no-nuts --no-other-workloads --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/ | tail -5
