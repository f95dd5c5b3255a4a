#!/bin/bash
# final state of round 2: GPU suite, the driver's bench command, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]); print("clocks", d["clocks"])
print("nuts", {k: d["nuts"].get(k) for k in ("ess_min_per_s", "wall_s", "rhat_max", "useful_eval_frac")} if "nuts" in d else None)
print("small", {k: (v["us_per_eval"], v["roofline"]["frac"], v["roofline"]["frac_algorithmic"]) for k, v in d.get("small_batch", {}).items() if k.startswith("c")})
print("strict", d.get("strict_math")); print("other", {k: (v.get("ms_per_step"), v.get("value")) for k, v in d.get("other_workloads", {}).items()})
print("cpu", d.get("cpu_baseline"))
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 120 --csv --log-file gpurun_out/r02_launches_final2.csv python bench.py --steps 5 --warmup 3 --no-nuts --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"
