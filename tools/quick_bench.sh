#!/bin/bash
# tests + bench summary in one gpurun call
timeout 300 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/gpu_tests.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/gpu_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.3e obj/s  ms/step %.3f  step_frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["step_frac"]))
for k,v in d["roofline"]["kernels"].items(): print("  %-22s %.3f ms  %.0f GB/s" % (k, v["ms"], v["gbs"]))
for k,v in d["configs"].items(): print("  %-32s %.4f ms  %.3e obj/s  frac %.3f" % (k, v["ms"], v["objects_per_s"], v["frac"]))
print("  e2e", d["e2e"]["value"], "clocks", d["clocks"])
PY
