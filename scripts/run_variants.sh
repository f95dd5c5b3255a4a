for v in ${VARIANTS:-0 1}; do BL_CHAIN_VARIANT=$v python bench.py --no-cpu-baseline --no-nuts --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', round(d['value']), round(d['ms_per_step'],3))"; done
