#!/usr/bin/env python
"""Tiny driver for ncu captures: runs one BASELINE config a few times.  tools/prof_case.py c2|c3|c4|c5"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')

case = sys.argv[1] if len(sys.argv) > 1 else 'c2'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda', 0)
kinv = pf.default_kinv(dev)
if case in ('c2', 'c3'):
    d = pf.synth.make_objects(4096, 64, 64, seed=2000, device=dev, n_hyp=128)
elif case == 'c4':
    d = pf.synth.make_objects(384, 112, 112, seed=4000, device=dev)
else:
    d = pf.synth.make_objects(32768, 64, 64, seed=5000, device=dev)
b = d['noc'].shape[0]
g = (torch.randn(b, device=dev), torch.randn(b, 9, device=dev), torch.randn(b, 3, device=dev))
torch.cuda.synchronize()
for _ in range(reps):
    raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv,
                          sample_idx=d['sample_idx'] if case == 'c3' else None)
    if case in ('c4', 'c5'):
        pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status, *g)
torch.cuda.synchronize()
print('ok', case, int((raw.status == 0).sum()), 'of', b)
