# N-GPU C2 bench (NGPU=2|4|8, default 2) under environment-knob sets ("A=1" = defaults; join knobs of one run with commas): value, ms/update, all-reduce mode
run() { echo "== $*"; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2} --steps ${STEPS:-5} --warmup ${WARM:-3} --no-large --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), (d.get('allreduce') or '')[:9], d['final_metrics']['critic_loss'], d.get('mc_phases') or '')"; }
for a in "$@"; do run $(echo $a | tr ',' ' '); done
