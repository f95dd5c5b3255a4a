#!/bin/bash
Q="--no-nuts --no-other-workloads --no-cpu-baseline --steps 10"
for v in 0 2 11 12 21; do
    BL_SIGNED_NS=$v python bench.py $Q --theta mode > gpurun_out/r02_k1d_g_$v.json 2>gpurun_out/r02_k1d_g.err
    python -c "
import json
d=json.load(open('gpurun_out/r02_k1d_g_$v.json')); print('variant $v mode', round(d['ms_per_step'],3), round(d['value']), d['clocks']['sm_mhz'])"
done
