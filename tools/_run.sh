export PROBE_MODES=rotate POSEFIT_DEPTH=4
for ed in 0 1 3 7 15 11 9; do echo "EARLY_DEP=$ed"; POSEFIT_EARLY_DEP=$ed python tools/latency_probe.py | grep rotate; done
unset POSEFIT_DEPTH
for ed in 1 15 9; do POSEFIT_EARLY_DEP=$ed python bench.py --steps 30 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ED=$ed', 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'])"; done
