python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/s16.json 2> gpurun_out/s16.err; tail -3 gpurun_out/s16.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s16.json").read().strip().splitlines()[-1]); print(d["e2e"], d["ms_per_step"])
PY
