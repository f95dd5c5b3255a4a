#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_small2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_small2_tests.log
timeout 600 python scripts/small_kernel_sweep.py > gpurun_out/r02_small_kernel_sweep2.txt 2>&1; cat gpurun_out/r02_small_kernel_sweep2.txt
BL_HIER_REDUCE=0 timeout 300 python scripts/small_kernel_sweep.py 1000000 quick > gpurun_out/r02_small_kernel_sweep2_nohier.txt 2>&1; cat gpurun_out/r02_small_kernel_sweep2_nohier.txt
