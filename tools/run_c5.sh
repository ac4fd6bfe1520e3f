# C5 (wide-net / 256k-sample) update under different kernel selections: samples/s, ms per update, launches per update
run() { echo "== $*"; env "$@" timeout 300 python bench.py --config C5 --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['launches_per_update'], round(d['roofline']['achieved'],1), round(d['roofline'].get('in_update_tflops',0),1))"; }
for a in "$@"; do run $a; done
