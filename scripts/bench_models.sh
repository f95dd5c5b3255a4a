for w in occu_rn_200k_x10_k50 occu_cop_500k_x12; do for c in 64 256 1024; do python bench.py --workload $w --chains $c --steps 5 --no-nuts --no-cpu-baseline 2>&1 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$w C=$c', round(d['value']), 'evals/s', round(d['ms_per_step'],2),'ms')
    elif ln.strip(): print(ln.strip()[:200])"; done; done
python bench.py --strict-math --steps 5 --no-nuts --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occu strict', round(d['value']), round(d['ms_per_step'],2))"
python bench.py --chains 8 --steps 10 --no-nuts --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occu C=8 (engine)', round(d['value']), round(d['ms_per_step'],3))"
python bench.py --chains 1 --steps 10 --no-nuts --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('occu C=1 (engine)', round(d['value']), round(d['ms_per_step'],3))"
