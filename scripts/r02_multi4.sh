#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_multi4_tests.log 2>&1; tail -3 gpurun_out/r02_multi4_tests.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err ) 2>&1 | grep real
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench_n4.err | tail -5
python - <<'PY'
import json
txt=open('gpurun_out/r02_bench_n4.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'])
print(json.dumps(d['site_sharded'])[:3000])
PY
