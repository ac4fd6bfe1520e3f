# C2 update under knob sets given as arguments ("A=1" = defaults; join several knobs of one run with commas)
run() { echo "== $*"; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-large --no-extra --steps 3 2>/tmp/run_knobs.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['launches_per_update'])" || tail -30 /tmp/run_knobs.err; }
for a in "$@"; do run $(echo $a | tr ',' ' '); done
