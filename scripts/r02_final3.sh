#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/small_kernel_sweep.py 1000000 quick 2>&1 | tee gpurun_out/r02_small_kernel_sweep_final2.txt | head -3
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-nuts > gpurun_out/r02_bench_n1_final_nonuts.json 2> gpurun_out/r02_bench_n1_final_nonuts.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final_nonuts.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "roofline frac", d["roofline"]["frac"])
print("small", {k: (v["us_per_eval"], v["us_per_eval_single_launch_after_l2_flush"], v["roofline"]["frac"], v["roofline"]["frac_algorithmic"]) for k, v in d.get("small_batch", {}).items() if k.startswith("c")})
print("strict", d.get("strict_math", {}).get("ms_per_step")); print("other", {k: (v.get("ms_per_step"), v.get("value")) for k, v in d.get("other_workloads", {}).items()})
P
