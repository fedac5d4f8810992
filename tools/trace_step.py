#!/usr/bin/env python
"""Timeline of one forward + backward step of the plain path from the PF_TRACE debug build (tooling, like tests/).

Build:  nvcc ... -DPF_TRACE -shared csrc/posefit_kernels.cu -o tools/_dbg/libposefit_trace.so
Run:    POSEFIT_LIB=tools/_dbg/libposefit_trace.so python tools/trace_step.py [--objects 384 --size 112] [--reps 6]

Every kernel stamps %globaltimer at its first instruction, after its griddepcontrol.wait and at its end (earliest /
latest over all warps).  The step is replayed as a CUDA graph rotating over input sets that exceed L2, like bench.py's
BASELINE configs 2 and 4; the stamps of one replay on an idle GPU (the counters can only be reset between replays) are
printed relative to the first instruction of its streaming forward kernel -- the host's launch latency is not in them."""
import argparse
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')

NAMES = {0: 'moments', 1: 'solve', 2: 'coefficients', 3: 'backward'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--objects', type=int, default=384)
    ap.add_argument('--size', type=int, default=112)
    ap.add_argument('--sets', type=int, default=8)
    ap.add_argument('--reps', type=int, default=6)
    ap.add_argument('--forward-only', action='store_true')
    a = ap.parse_args()
    lib = pf._lib.lib()
    if not hasattr(lib, 'posefit_debug_trace'):
        raise SystemExit('this library was not built with -DPF_TRACE (set POSEFIT_LIB)')
    lib.posefit_debug_trace.restype = ctypes.c_int
    lib.posefit_debug_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    dev = torch.device('cuda')
    kinv = pf.default_kinv(dev)
    n, sz = a.objects, a.size
    sets = [pf.synth.make_objects(n, sz, sz, seed=4000 + i, device=dev) for i in range(a.sets)]
    g = (torch.randn(n, device=dev), torch.randn(n, 9, device=dev), torch.randn(n, 3, device=dev))

    def step(c):
        raw = pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv)
        if not a.forward_only:
            pf.pose_fit_backward_raw(c['noc'], c['depth'], c['mask'], None, c['bbox_xy0'], kinv, raw.ctx, raw.status, *g)

    graphs = []
    for c in sets:
        for _ in range(2):
            step(c)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(c)
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                step(c)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs.append(graph)
    buf = (ctypes.c_ulonglong * 32)()
    rows = []
    for rep in range(a.reps):
        for i in range(3):                                       # warm replays (instruction caches, clocks)
            graphs[(rep + i) % len(graphs)].replay()
        torch.cuda.synchronize()
        lib.posefit_debug_trace(buf, 1)                          # reset the stamps (needs an idle GPU)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        graphs[(rep + 3) % len(graphs)].replay()
        a1.record()
        torch.cuda.synchronize()
        lib.posefit_debug_trace(buf, 0)
        t = list(buf)
        wbuf = (ctypes.c_ulonglong * 4096)()
        if hasattr(lib, 'posefit_debug_trace_warps'):
            lib.posefit_debug_trace_warps(wbuf)
        rows.append((a0.elapsed_time(a1) * 1e3, t, list(wbuf)))
    rows.sort(key=lambda r: r[0])
    us, t, wt = rows[len(rows) // 2]                             # the median replay
    t0 = t[0]
    print(f'{n} objects {sz}x{sz}, one replay on an idle GPU: {us:.1f} us between the events around it; stamps relative '
          f'to the first instruction of the streaming forward kernel')
    for k in (0, 1, 2, 3):
        if t[2 * k] == 2 ** 64 - 1:
            continue
        print(f'{NAMES[k]:13s} first instruction {(t[2 * k] - t0) / 1e3:7.2f}  past its wait {(t[2 * (8 + k)] - t0) / 1e3:7.2f}  '
              f'last warp done {(t[2 * k + 1] - t0) / 1e3:7.2f} us')
        if k == 0:
            warp_report(wt, t0)
        if k in (0, 3):                                          # streaming kernels: how far apart their warps start / end
            a_, b_ = (4, 5) if k == 0 else (6, 7)
            print(f'    latest first instruction {(t[2 * a_ + 1] - t0) / 1e3:7.2f}, earliest warp done {(t[2 * b_] - t0) / 1e3:7.2f} us')
        if k == 1:
            def span(i):
                return f'{(t[2 * i] - t0) / 1e3:.2f} .. {(t[2 * i + 1] - t0) / 1e3:.2f}'
            print(f'    solve, earliest .. latest warp: past the wait {span(9)}, moments merged {span(12)}, solved {span(13)} us')


def warp_report(wt, t0):
    """moments kernel: when the warps of every CTA finished (g_trace_warp) -- spread inside a CTA vs between CTAs"""
    import numpy as np
    w = (np.array(wt, dtype=np.float64).reshape(256, 16) - t0) / 1e3
    live = w[(w > 0).all(axis=1) & (w < 1e6).all(axis=1)]
    if len(live) == 0:
        return
    print(f'    warps finished (us after the first instruction), {len(live)} CTAs x 16 warps: all {live.min():.1f} .. {live.max():.1f}; '
          f'per CTA first {live.min(axis=1).mean():.1f} / last {live.max(axis=1).mean():.1f} on average; CTA means '
          f'{live.mean(axis=1).min():.1f} .. {live.mean(axis=1).max():.1f}')
    print('    by warp index (mean over CTAs): ' + ' '.join(f'{x:.1f}' for x in live.mean(axis=0)))
    if os.environ.get('TRACE_PER_CTA'):
        print('    last warp of every CTA, in blockIdx order: ' + ' '.join(f'{x:.0f}' for x in live.max(axis=1)))


if __name__ == '__main__':
    main()
