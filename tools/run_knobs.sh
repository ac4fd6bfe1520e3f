run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['launches_per_update'])"; }
run A=1
run REPPO_TC_WIDE_MIN=60
run REPPO_NORM_BN128=1
run REPPO_TC_WIDE_MIN=60 REPPO_NORM_BN128=1
run REPPO_WGRAD_SPLIT=2
run REPPO_SLAB=1
