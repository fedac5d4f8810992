#!/usr/bin/env python
"""Small invocation of every C-ABI entry (for compute-sanitizer --tool memcheck)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
for (h, w, b, nh) in ((64, 64, 300, 128), (112, 112, 40, 0), (19, 27, 9, 12), (32, 48, 20, 24)):
    d = pf.synth.make_objects(b, h, w, seed=1, device='cuda', n_hyp=max(nh, 4), align_x0=1 if w % 4 else 4)
    noc = d['noc'].clone().requires_grad_(True)
    out = pf.pose_fit(noc, d['depth'], d['mask'], d['bbox_xy0'])
    (out[0].sum() + out[1].sum() + out[2].sum()).backward()
    raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'])
    if nh:
        noc2 = d['noc'].clone().requires_grad_(True)
        o2 = pf.pose_fit(noc2, d['depth'], d['mask'], d['bbox_xy0'], sample_idx=d['sample_idx'])
        (o2[0].sum() + o2[1].sum() + o2[2].sum()).backward()
    cam = torch.eye(4, dtype=torch.float64)
    pf.pose_epilogue(raw, d['depth'], d['mask'], d['bbox_xy0'], campose=cam)
    box = torch.randn(b, 8, 3, dtype=torch.float64)
    pf.clip_mask_to_box(d['depth'], d['mask'], d['bbox_xy0'], box, cam)
    if b <= 20:
        m1 = pf.statistical_outlier_mask(None, d['depth'], d['mask'], d['bbox_xy0'])
        pf.statistical_outlier_mask(d['noc'], d['depth'], m1, d['bbox_xy0'], source='noc')
src = torch.randn(5, 3, 700, dtype=torch.float64, device='cuda')
dst = torch.randn(5, 3, 700, dtype=torch.float64, device='cuda')
idx = torch.randint(0, 700, (5, 33, 10), dtype=torch.int32, device='cuda')
pf.points_fit_raw(src, dst)
pf.points_fit_raw(src, dst, sample_idx=idx)
torch.cuda.synchronize()
print('sanitize case done')
