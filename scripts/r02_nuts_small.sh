#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nuts.py -m gpu -x -q 2>&1 | tail -3
for st in 1 0; do
  echo "BL_NUTS_STAGED=$st, 5 chains, 1000 + 1000, 1M sites:"; BL_NUTS_STAGED=$st timeout 300 python scripts/nuts_probe.py --chains 5 --warmup 1000 --samples 1000 2>&1 | tail -1
done | tee gpurun_out/r02_nuts_5chains_staged.txt
