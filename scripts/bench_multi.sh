#!/bin/bash
# usage: scripts/bench_multi.sh N   (torchrun, one rank per GPU)
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $N "$@" 2>&1 | grep -E '^\{|Error|error' ; }
echo "== chain-sharded occu_1m_x8_c1024"; run --steps 10 --no-cpu-baseline
echo "== site-sharded p2p"; run --workload occu_sites16m_c256 --exchange p2p --steps 20 --no-cpu-baseline --nuts-warmup 100 --nuts-samples 50
echo "== site-sharded nccl"; run --workload occu_sites16m_c256 --exchange nccl --steps 20 --no-nuts --no-cpu-baseline
