#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py into profiles/rNN_bench_launches.md.
usage: tools/launch_list.py <launches.csv> <out.md> [objects_per_step_kernel_group=4]"""
import csv, io, sys
from collections import OrderedDict

src, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
launches = [(r['Kernel Name'], float(r['Metric Value']) / 1e3) for r in rows if r['Metric Name'] == 'gpu__time_duration.sum']
# the headline step = the first repeating group of 4 launches (moments, solve, coef, backward) at the large size
head = launches[:32]
per = OrderedDict()
for name, us in head:
    per.setdefault(name, []).append(us)
step = sum(sum(v) / len(v) for v in per.values())
with open(out, 'w') as f:
    f.write('# ncu launch list of `python bench.py --steps 5 --warmup 3 --no-cpu` (r02, final kernels)\n\n')
    f.write('Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fit_|pose_" -c 400 --csv python bench.py --steps 5 --warmup 3 --no-cpu`\n\n')
    f.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes.\n\n')
    f.write('Headline step (config-5 shard, mean of the first 8 steps):\n\n| kernel | us per step | share |\n|---|---|---|\n')
    for name, v in per.items():
        m = sum(v) / len(v)
        f.write(f'| {name.replace("void ", "").replace("(FwdParams)", "").replace("(BwdParams)", "")} | {m:.1f} | {100 * m / step:.1f}% |\n')
    f.write('\nAll launches:\n\n| # | kernel | duration (us) |\n|---|---|---|\n')
    for i, (name, us) in enumerate(launches):
        f.write(f'| {i} | {name} | {us:.1f} |\n')
print('step under ncu: %.1f us over %d launches' % (step, len(launches)))
