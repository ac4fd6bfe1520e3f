# 2-GPU C2 bench under knob sets (same syntax as run_knobs.sh)
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-large 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))"; }
for a in "$@"; do run $(echo $a | tr ',' ' '); done
