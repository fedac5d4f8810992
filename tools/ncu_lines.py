#!/usr/bin/env python
"""Aggregate an .ncu-rep's SASS profile by CUDA source line (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
per = defaultdict(lambda: [0, 0, ''])      # (file, line) -> [inst, samples, text]
fname = ''
hdr = None
cur = None
for r in rows:
    if len(r) == 2 and r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]
        continue
    if len(r) > 4 and r[0] == 'Line No':
        hdr = r
        iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
        continue
    if hdr is None or len(r) <= iex:
        continue
    if r[0]:                                # a CUDA line row carries the totals of its SASS rows
        cur = (fname, int(r[0]))
        per[cur][2] = r[1].strip()
        per[cur][0] += int(r[iex]) if r[iex].isdigit() else 0
        per[cur][1] += int(r[ismp]) if r[ismp].isdigit() else 0
ti = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values())
print(f'total warp-instructions {ti}, samples {ts}')
for (f, l), (i, s, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{f}:{l:5d} inst {100*i/ti:5.1f}% samp {100*s/ts:5.1f}%  {t[:110]}')
