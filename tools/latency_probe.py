#!/usr/bin/env python
"""How the small configs (C2 / C4) respond to the way the L2 is flushed between CUDA-graph replays.
tools/latency_probe.py  ->  min / median device time per methodology."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
dev = torch.device('cuda', 0)
kinv = pf.default_kinv(dev)
flush_w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_r = torch.ones(64 << 20, dtype=torch.int32, device=dev)
sink = torch.zeros((), dtype=torch.int64, device=dev)


def graph_of(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    return g


def timeit(graphs, mode, reps=40):
    tt = []
    for i in range(reps):
        g = graphs[i % len(graphs)]
        if mode in ('dirty', 'clean'):
            flush_w.zero_()
        if mode == 'clean':
            sink.copy_(flush_r.sum())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        tt.append(a.elapsed_time(b) * 1e3)
    tt.sort()
    return tt[0], tt[len(tt) // 2]


def run(name, make, step, n_sets):
    sets = [make(s) for s in range(n_sets)]
    graphs = [graph_of(lambda d=d: step(d)) for d in sets]
    modes = os.environ.get('PROBE_MODES', 'dirty,clean,rotate,same').split(',')
    for mode, gs in (('dirty', graphs[:1]), ('clean', graphs[:1]), ('rotate', graphs), ('same', graphs[:1])):
        if mode not in modes:
            continue
        mn, md = timeit(gs, mode)
        print(f'{name:4s} {mode:7s} min {mn:7.1f} us  median {md:7.1f} us')


def c2_make(s):
    return pf.synth.make_objects(4096, 64, 64, seed=2000 + s, device=dev)


def c2_step(d):
    pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)


def c4_make(s):
    d = pf.synth.make_objects(384, 112, 112, seed=4000 + s, device=dev)
    d['g'] = (torch.randn(384, device=dev), torch.randn(384, 9, device=dev), torch.randn(384, 3, device=dev))
    return d


def c4_step(d):
    raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)
    pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status, *d['g'])


print('POSEFIT_PREWARM =', os.environ.get('POSEFIT_PREWARM', 'auto'))
run('C2', c2_make, c2_step, 4)
run('C4', c4_make, c4_step, 8)
