for mode in sync async sync async; do
POSEFIT_BENCH_GATHER=$mode timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --no-extra --e2e-objects 256 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', 'ms/step %.3f'%d['ms_per_step'], 'value %.3e'%d['value'], d['clocks']['sm_mhz'])"
done
