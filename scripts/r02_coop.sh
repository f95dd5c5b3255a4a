for v in 0 1 0 1; do echo "BL_COOP_REDUCE=$v"; BL_COOP_REDUCE=$v python scripts/lane_sweep.py 2>&1 | tail -1; done
python bench.py --no-nuts --no-other-workloads --no-cpu-baseline --steps 10 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['clocks'])"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
