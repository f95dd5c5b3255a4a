#!/bin/bash
# K1s (small chain batches): parity tests, sweep against the engine + oracle, ncu captures at C = 1 and C = 5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_small.py -m gpu -x -q > gpurun_out/r02_small_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_small_tests.log
timeout 600 python scripts/small_kernel_sweep.py > gpurun_out/r02_small_kernel_sweep_final.txt 2>&1; cat gpurun_out/r02_small_kernel_sweep_final.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:occu_small_kernel -s 3 -c 1 -o /tmp/k1s_c1 python scripts/small_batch_probe.py > gpurun_out/r02_ncu_s1.log 2>&1
python profiles/ncu_summary.py /tmp/k1s_c1.ncu-rep $((1000000*1/32)) gpurun_out/r02_occu_small_c1.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:occu_small_kernel -s 7 -c 1 -o /tmp/k1s_c5 python scripts/small_batch_probe.py > gpurun_out/r02_ncu_s5.log 2>&1
python profiles/ncu_summary.py /tmp/k1s_c5.ncu-rep $((1000000*5/32)) gpurun_out/r02_occu_small_c5.txt
head -34 gpurun_out/r02_occu_small_c1.txt; head -34 gpurun_out/r02_occu_small_c5.txt
