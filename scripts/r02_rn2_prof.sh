#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_full_tests.log 2>&1; tail -3 gpurun_out/r02_full_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:occu_rn2_kernel -s 3 -c 1 -o gpurun_out/r02_rn2 python bench.py --workload occu_rn_200k_x10_k50 --steps 1 --warmup 3 --no-nuts --no-cpu-baseline > gpurun_out/r02_rn2_ncu.log 2>&1
ls -la gpurun_out/r02_rn2.ncu-rep
