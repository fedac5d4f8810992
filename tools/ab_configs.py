#!/usr/bin/env python
"""A/B timing of the small BASELINE configs under different launch knobs, in one process (tooling, like tests/).

    python tools/ab_configs.py [--configs c2,c3,c4] [--reps 40] VARIANT [VARIANT ...]

A VARIANT is a comma-separated list of POSEFIT_* assignments without the prefix, e.g. `RANSAC_THREADS=192` or
`RANSAC_SCREEN=0,RANSAC_THREADS=256`; `base` is the library's defaults.  Every config is timed the way bench.py
times it (CUDA-graph replays rotating over input sets that together exceed L2, queued back to back; the isolated-replay median beside it), and for C3 the
inlier masks / winners / poses of every variant are compared with those of the first variant."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')


def timed(fns, reps):
    graphs = []
    for fn in fns:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs.append(graph)
    tt = []
    for i in range(reps + len(graphs)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graphs[i % len(graphs)].replay()
        b.record()
        torch.cuda.synchronize()
        if i >= len(graphs):
            tt.append(a.elapsed_time(b))
    tt.sort()
    bb = []
    for _ in range(5):                                           # back to back, as bench.py reports it
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graphs[-1].replay()
        a.record()
        for i in range(reps):
            graphs[i % len(graphs)].replay()
        b.record()
        torch.cuda.synchronize()
        bb.append(a.elapsed_time(b) / reps)
    bb.sort()
    return bb[len(bb) // 2], tt[len(tt) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='c2,c3,c4')
    ap.add_argument('--reps', type=int, default=40)
    ap.add_argument('variants', nargs='*', default=['base'])
    a = ap.parse_args()
    dev = torch.device('cuda')
    kinv = pf.default_kinv(dev)
    peak = 6552.6
    cfgs = a.configs.split(',')
    data = {}
    if 'c2' in cfgs or 'c3' in cfgs:
        data['c2'] = [pf.synth.make_objects(4096, 64, 64, seed=2000 + i, device=dev, n_hyp=128) for i in range(4)]
    if 'c4' in cfgs:
        data['c4'] = [pf.synth.make_objects(384, 112, 112, seed=4000 + i, device=dev) for i in range(8)]
        g4 = (torch.randn(384, device=dev), torch.randn(384, 9, device=dev), torch.randn(384, 3, device=dev))
    ref = None
    set_names = set()
    for var in a.variants:
        for name in set_names:
            os.environ.pop(name, None)
        set_names = set()
        if var != 'base':
            for kv in var.split(','):
                k, v = kv.split('=')
                os.environ['POSEFIT_' + k] = v
                set_names.add('POSEFIT_' + k)
        pf._lib.reload_knobs()
        line = [f'{var:40s}']
        if 'c2' in cfgs:
            ms, lo = timed([lambda c=c: pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv)
                            for c in data['c2']], a.reps)
            b2 = 4096 * (17 * 4096 + 64)
            line.append(f'C2 {ms * 1e3:7.1f} us ({b2 / ms / 1e6 / peak:.3f}) isolated {lo * 1e3:6.1f}')
        if 'c3' in cfgs:
            ms, lo = timed([lambda c=c: pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv,
                                                        sample_idx=c['sample_idx']) for c in data['c2']], a.reps)
            b3 = 4096 * (17 * 4096 + 64 + 128 * 10 * 4 + 4096)
            c = data['c2'][0]
            out = pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv, sample_idx=c['sample_idx'])
            torch.cuda.synchronize()
            same = ''
            if ref is None:
                ref = out
            else:
                ok = (torch.equal(out.inlier_mask, ref.inlier_mask) and torch.equal(out.winner, ref.winner)
                      and torch.equal(out.status, ref.status))
                dp = float((out.pose[:, :13] - ref.pose[:, :13]).abs().max())
                same = f' same={ok} dpose={dp:.1e}'
            line.append(f'C3 {ms * 1e3:7.1f} us ({b3 / ms / 1e6 / peak:.3f}) isolated {lo * 1e3:6.1f}{same}')
        if 'c4' in cfgs:
            def c4_step(c4):
                raw = pf.pose_fit_raw(c4['noc'], c4['depth'], c4['mask'], c4['bbox_xy0'], kinv)
                pf.pose_fit_backward_raw(c4['noc'], c4['depth'], c4['mask'], None, c4['bbox_xy0'], kinv, raw.ctx,
                                         raw.status, *g4)
            ms, lo = timed([lambda c=c: c4_step(c) for c in data['c4']], a.reps)
            b4 = 384 * (46 * 112 * 112)
            line.append(f'C4 {ms * 1e3:7.1f} us ({b4 / ms / 1e6 / peak:.3f}) isolated {lo * 1e3:6.1f}')
        print(' | '.join(line), flush=True)


if __name__ == '__main__':
    main()
