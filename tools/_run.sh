python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/s5_bench.json 2>gpurun_out/s5_bench.err; python - <<'PY'
import json,sys
d=json.loads(open("gpurun_out/s5_bench.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]))
for k,v in d["configs"].items(): print("  %-32s %.4f ms frac %.3f" % (k, v["ms"], v["frac"]))
PY
