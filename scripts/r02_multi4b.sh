#!/bin/bash
mkdir -p gpurun_out
( time timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --no-nuts --steps 5 --site-sharded-timeout 150 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err ) 2>&1 | grep real
grep "site_sharded\|Error\|error" gpurun_out/r02_bench_n4.err | tail -20
python - <<'PY'
import json
txt=open('gpurun_out/r02_bench_n4.json').read()
ls=[l for l in txt.splitlines() if l.startswith('{')]
if ls:
    d=json.loads(ls[-1]); print(d['n_gpus'], d['ms_per_step'], d['value']); print(json.dumps(d.get('site_sharded'))[:3500])
else: print("no json line")
PY
