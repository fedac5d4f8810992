#!/usr/bin/env python
"""Per-kernel digest of the built library's SASS (no GPU needed): registers / stack / shared memory from
`cuobjdump --dump-resource-usage`, instruction totals and the opcodes that prove what the kernel does
(UBLKCP / SYNCS = TMA bulk copies + mbarriers, LDGSTS = cp.async, DFMA / DADD / DMUL = double arithmetic, F2F = the
float -> double conversions, MUFU, LDL / STL = local memory).  usage: tools/sass_digest.py [lib.so] > profiles/rNN_sass_digest.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, '3d_mot_differentiable_pose_estimation_b200', 'libposefit_b200.so')
res = subprocess.run(['cuobjdump', '--dump-resource-usage', lib], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r'Function (\S+):', line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r'REG:(\d+).*STACK:(\d+).*SHARED:(\d+)', line)
    if m and name:
        usage[name] = tuple(int(v) for v in m.groups())
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        kern[cur][m.group(1)] += 1
KEYS = ['UBLKCP', 'SYNCS', 'LDGSTS', 'DFMA', 'DADD', 'DMUL', 'F2F', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'LDL', 'STL', 'BAR']
print('# SASS digest of ' + os.path.basename(lib) + ' (sm_100a; `tools/sass_digest.py`)\n')
print('No tensor-core opcodes (UTC*MMA / HMMA) anywhere: nothing on this path is a contraction.  UBLKCP + SYNCS = 1-D TMA bulk')
print('copies with mbarriers (the RANSAC kernels stage whole crops), LDGSTS = cp.async (per-warp rings of the moments kernel).\n')
print('| kernel | regs | stack B | static smem B | instr | ' + ' | '.join(KEYS) + ' |')
print('|---|---|---|---|---|' + '---|' * len(KEYS))
for k, c in kern.items():
    short = subprocess.run(['c++filt', '-p', k], capture_output=True, text=True).stdout.strip() or k
    short = short.replace('posefit::', '')
    r = usage.get(k, ('?', '?', '?'))
    print(f'| `{short}` | {r[0]} | {r[1]} | {r[2]} | {sum(c.values())} | ' + ' | '.join(str(c.get(x, 0)) for x in KEYS) + ' |')
tc = sum(v for c in kern.values() for op, v in c.items() if op.startswith('UTC') or op.startswith('HMMA') or op.startswith('HGMMA'))
print(f'\ntensor-core instructions in the library: {tc}')
