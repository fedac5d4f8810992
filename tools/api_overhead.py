#!/usr/bin/env python
"""Where the public autograd operator's time goes against the raw C-ABI calls on the config-5 shard (tooling)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
dev = torch.device('cuda')
d = pf.synth.make_objects(n, 64, 64, seed=5000, device=dev)
kinv = pf.default_kinv(dev)
g = (torch.randn(n, device=dev), torch.randn(n, 3, 3, device=dev), torch.randn(n, 3, device=dev))


def timeit(fn, k=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


def raw_fwd():
    return pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)


def raw_both():
    raw = raw_fwd()
    pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status, g[0], g[1], g[2])


def api_fwd(mask=False):
    with torch.no_grad():
        return pf.pose_fit(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], return_mask=mask)


def api_both(mask=False):
    noc = d['noc'].requires_grad_(True)
    noc.grad = None
    s, r, t, _, _, _ = pf.pose_fit(noc, d['depth'], d['mask'], d['bbox_xy0'], return_mask=mask)
    torch.autograd.backward((s, r, t), g)


print('raw fwd            %.3f ms' % timeit(raw_fwd))
print('api fwd            %.3f ms' % timeit(api_fwd))
print('api fwd + mask     %.3f ms' % timeit(lambda: api_fwd(True)))
print('raw fwd + bwd      %.3f ms' % timeit(raw_both))
print('api fwd + bwd      %.3f ms' % timeit(api_both))
print('api fwd + bwd+mask %.3f ms' % timeit(lambda: api_both(True)))
