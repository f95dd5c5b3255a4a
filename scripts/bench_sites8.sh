#!/bin/bash
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N --workload occu_sites16m_c256 --exchange ${2:-p2p} --steps 20 --no-cpu-baseline --nuts-warmup 300 --nuts-samples 100 2>&1 | grep -E '^\{|Error|error'
