for c in 64 256 1024; do python bench.py --workload occu_cop_500k_x12 --chains $c --steps 5 --no-nuts --no-cpu-baseline 2>&1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('cop C=$c', round(d['value']), 'evals/s', round(d['ms_per_step'],2),'ms')
    elif ln.strip(): print(ln.strip()[:200])"; done
