#!/usr/bin/env python
"""Condense an .ncu-rep into the few numbers the roofline argument needs (+ the hottest SASS lines).
usage: tools/ncu_summary.py <report.ncu-rep> [out.md]"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes.sum.per_second',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
STALLS = 'smsp__average_warps_issue_stalled_'


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    raw = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        name = row[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        out.append(f'## {name}\n')
        out.append('| metric | value | unit |\n|---|---|---|')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f'| {k} | {row[i]} | {units[i]} |')
        stalls = sorted(((float(row[i]), h[len(STALLS):].replace('_per_issue_active.ratio', ''))
                         for i, h in enumerate(hdr) if h.startswith(STALLS) and h.endswith('per_issue_active.ratio')),
                        reverse=True)
        out.append('\nstall reasons (warps per issue-active cycle): ' +
                   ', '.join(f'{n} {v:.2f}' for v, n in stalls[:7]) + '\n')
    src = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    try:
        h = next(r for r in src if 'Address' in r and 'Source' in r)
        rows = src[src.index(h) + 1:]
        ia, isrc, iex, ismp = h.index('Address'), h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
        data = [(r[isrc], int(r[iex] or 0), int(r[ismp] or 0)) for r in rows if len(r) > iex]
        tot_i, tot_s = sum(d[1] for d in data), sum(d[2] for d in data)
        out.append(f'SASS (first kernel): {tot_i} warp-instructions, {tot_s} samples; instructions grouped by execution count:\n')
        out.append('| executions per instruction | #instructions | share of instructions | share of samples |\n|---|---|---|---|')
        ce, cs, cn = Counter(), Counter(), Counter()
        for s, e, m in data:
            ce[e] += e
            cs[e] += m
            cn[e] += 1
        for e, t in sorted(ce.items(), key=lambda x: -x[1])[:8]:
            out.append(f'| {e} | {cn[e]} | {100 * t / max(tot_i, 1):.1f}% | {100 * cs[e] / max(tot_s, 1):.1f}% |')
        ops = Counter()
        for s, e, m in data:
            op = s.split()[1] if s.startswith('@') and len(s.split()) > 1 else (s.split()[0] if s.split() else '')
            ops[op.split('.')[0]] += e
        out.append('\nexecuted opcodes: ' + ', '.join(f'{k} {100 * v / max(tot_i, 1):.1f}%' for k, v in ops.most_common(14)) + '\n')
    except StopIteration:
        pass
    text = '\n'.join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')
    print(text)


if __name__ == '__main__':
    main()
