#!/bin/bash
# final ncu captures of the two rewritten kernels, summarised ON the box (the .ncu-rep files exceed the 64 MiB return limit)
mkdir -p gpurun_out
Q="--steps 1 --warmup 3 --no-nuts --no-other-workloads --no-cpu-baseline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:occu_signed_kernel -s 3 -c 1 -o /tmp/k1d_final python bench.py $Q > gpurun_out/r02_ncu_a.log 2>&1
python profiles/ncu_summary.py /tmp/k1d_final.ncu-rep $((1000000*1024/32)) gpurun_out/r02_occu_signed_final.txt
timeout 500 ncu --set full --clock-control none --import-source on -k regex:occu_rn2_kernel -s 3 -c 1 -o /tmp/rn2_final python bench.py --workload occu_rn_200k_x10_k50 $Q > gpurun_out/r02_ncu_b.log 2>&1
python profiles/ncu_summary.py /tmp/rn2_final.ncu-rep $((200000*256/32)) gpurun_out/r02_occu_rn2_final.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 80 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 5 --warmup 3 --no-nuts --no-cpu-baseline > /dev/null 2>&1
python scripts/numerics_table.py > gpurun_out/r02_numerics_final.txt 2>&1; tail -2 gpurun_out/r02_numerics_final.txt
python scripts/mufu_error.py > gpurun_out/r02_mufu_error.txt 2>&1
python scripts/lane_sweep.py > gpurun_out/r02_small_batch.txt 2>&1
head -12 gpurun_out/r02_occu_signed_final.txt | tail -8
