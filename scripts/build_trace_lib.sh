#!/bin/bash
# libbiolith_b200_trace.so = the library with K1s' %globaltimer block stamps compiled in (-DBL_TRACE, occu_small.cu only);
# scripts/small_trace.py loads it instead of the product library.  Run after `python -m biolith_b200.build`.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-fvisibility=hidden --expt-relaxed-constexpr"
nvcc $F -DBL_TRACE -c biolith_b200/csrc/occu_small.cu -o /tmp/occu_small_trace.o
objs=$(ls biolith_b200/_build/*.o | grep -v "/occu_small.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o biolith_b200/libbiolith_b200_trace.so $objs /tmp/occu_small_trace.o -ldl
echo "built biolith_b200/libbiolith_b200_trace.so"
