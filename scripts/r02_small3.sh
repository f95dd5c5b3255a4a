#!/bin/bash
# K1s after the guard removal + what `fit` with the reference's default 5 chains costs end to end (device NUTS, 1M sites)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small.py tests/test_gpu_nuts.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/small_kernel_sweep.py 1000000 quick 2>&1 | tee gpurun_out/r02_small_kernel_sweep6.txt
for k in 1 0; do
  echo "BL_SMALL_KERNEL=$k, 5 chains, 1000 + 1000:"; BL_SMALL_KERNEL=$k timeout 300 python scripts/nuts_probe.py --chains 5 --warmup 1000 --samples 1000 2>&1 | tail -1
done | tee gpurun_out/r02_nuts_5chains.txt
