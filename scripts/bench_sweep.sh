set -x
python bench.py --workload occu_small --steps 5 --no-cpu-baseline --nuts-warmup 50 --nuts-samples 30 | python scripts/summarize_bench.py /dev/stdin
python bench.py --workload occu_small --dtype float64 --steps 3 --no-cpu-baseline --no-nuts | python scripts/summarize_bench.py /dev/stdin
python bench.py --workload occu_sites16m_c256 --steps 3 --no-cpu-baseline --no-nuts | python scripts/summarize_bench.py /dev/stdin
python bench.py --chains 1000 --steps 3 --no-cpu-baseline --no-nuts | python scripts/summarize_bench.py /dev/stdin
python bench.py --chains 100 --steps 3 --no-cpu-baseline --no-nuts | python scripts/summarize_bench.py /dev/stdin
python bench.py --workload occu_cop_500k_x12 --steps 3 --no-cpu-baseline --nuts-warmup 60 --nuts-samples 30 | python scripts/summarize_bench.py /dev/stdin
python bench.py --workload occu_rn_200k_x10_k50 --steps 2 --no-cpu-baseline --nuts-warmup 40 --nuts-samples 20 | python scripts/summarize_bench.py /dev/stdin
