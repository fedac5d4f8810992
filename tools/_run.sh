python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s8_bench.json 2>gpurun_out/s8_bench.err; python - <<'PY'
import json,sys
d=json.loads(open("gpurun_out/s8_bench.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
for k,v in d["configs"].items(): print("  %-32s %.4f ms frac %.3f" % (k, v["ms"], v["frac"]))
PY
