#!/bin/bash
# K1d (signed records) vs K1c A/B on config 2 + parity tests + one ncu capture
set -x
mkdir -p gpurun_out
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_xla_boundary.py -m gpu -q -x > gpurun_out/r02_k1d_tests.log 2>&1; tail -3 gpurun_out/r02_k1d_tests.log
BL_OCCU_CHAIN_KERNEL=1 python bench.py $Q > gpurun_out/r02_k1c.json 2>gpurun_out/r02_k1c.err
python bench.py $Q > gpurun_out/r02_k1d_ns2.json 2>gpurun_out/r02_k1d.err
BL_SIGNED_NS=1 python bench.py $Q > gpurun_out/r02_k1d_ns1.json 2>>gpurun_out/r02_k1d.err
BL_SIGNED_BT=128 python bench.py $Q > gpurun_out/r02_k1d_bt128.json 2>>gpurun_out/r02_k1d.err
python bench.py $Q --theta mode > gpurun_out/r02_k1d_mode.json 2>>gpurun_out/r02_k1d.err
for f in k1c k1d_ns2 k1d_ns1 k1d_bt128 k1d_mode; do python -c "
import json,sys
d=json.load(open('gpurun_out/r02_$f.json')); print('$f', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:occu_signed_kernel -s 3 -c 1 -o gpurun_out/r02_k1d python bench.py --steps 1 --warmup 3 --no-nuts --no-other-workloads --no-cpu-baseline > gpurun_out/r02_k1d_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
