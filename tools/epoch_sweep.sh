for e in 1 2 4 8; do timeout 200 python bench.py --no-cpu-baseline --no-large --steps 3 --override hyperparameters.num_epochs=$e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print($e, round(d['ms_per_step'],3), d['launches_per_update'])"; done
