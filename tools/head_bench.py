#!/usr/bin/env python
"""Head-fed fit (fused roi_align resize) against the composition it replaces, on the config-5 shard (tooling)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device('cuda')
d = pf.synth.make_objects(n, size, size, seed=5000, device=dev)
head0 = torch.nn.functional.adaptive_avg_pool2d(d['noc'], 28).contiguous()
del d['noc']
roi = torch.tensor([[size, size]], dtype=torch.int32, device=dev).repeat(n, 1)
g = (torch.randn(n, device=dev), torch.randn(n, 3, 3, device=dev), torch.randn(n, 3, device=dev))


def timeit(fn, k=10):
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


def composed(bwd=True):
    head = head0.requires_grad_(bwd)
    head.grad = None
    noc = pf.resample_noc(head, roi, size, size)
    s, r, t, _, _, _ = pf.pose_fit(noc, d['depth'], d['mask'], d['bbox_xy0'])
    if bwd:
        torch.autograd.backward((s, r, t), g)


def fused(bwd=True):
    head = head0.requires_grad_(bwd)
    head.grad = None
    s, r, t, _, _ = pf.pose_fit_head(head, roi, d['depth'], d['mask'], d['bbox_xy0'])
    if bwd:
        torch.autograd.backward((s, r, t), g)


P = size * size
print('%d objects %dx%d' % (n, size, size))
for name, fn in (('composed fwd', lambda: composed(False)), ('fused fwd', lambda: fused(False)),
                 ('composed fwd+bwd', composed), ('fused fwd+bwd', fused)):
    ms = timeit(fn)
    print('%-18s %.3f ms  %.2f M obj/s' % (name, ms, n / ms / 1e3))
