for c in 128 256 1024; do python bench.py --workload occu_rn_200k_x10_k50 --chains $c --steps 5 --no-nuts --no-cpu-baseline 2>&1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('rn C=$c', round(d['value']), 'evals/s', round(d['ms_per_step'],2),'ms')
    elif ln.strip(): print(ln.strip()[:200])"; done
