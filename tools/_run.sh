python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/s13.json 2> gpurun_out/s13.err; tail -3 gpurun_out/s13.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s13.json").read().strip().splitlines()[-1]); print(d["clocks"], d["ms_per_step"])
PY
