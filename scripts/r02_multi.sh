#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
( time python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02_bench_n1.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r02_bench_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference > gpurun_out/r02_bench_n2_ref.json 2> gpurun_out/r02_bench_n2_ref.err ) 2>&1 | grep real
python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_multi_tests.log 2>&1; tail -3 gpurun_out/r02_multi_tests.log
head -c 400 gpurun_out/r02_bench_n2_ref.json
