#!/usr/bin/env python
"""Per-phase cycle breakdown of fit_ransac_kernel (thread 0 of every CTA, summed over objects) from a debug build:
  make -C 3d_mot_differentiable_pose_estimation_b200/csrc EXTRA=-DPF_RANSAC_TIMING OUT=$PWD/gpurun_out/libposefit_timing.so
  python tools/ransac_phases.py gpurun_out/libposefit_timing.so"""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
libmod = importlib.import_module('3d_mot_differentiable_pose_estimation_b200._lib')
libmod.LIB_PATH = os.path.abspath(sys.argv[1])
lib = libmod.lib()
names = ['load wait + geometry', 'pass 1 loop', 'pass 1 reduction', 'centring / thresholds / prefix', 'select list',
         'sample gathers', 'hypothesis fit + residual', 'wait for other warps', 'selection + broadcast', 'pass 2 loop',
         'pass 2 reduction + record']
b = 4096
d = pf.synth.make_objects(b, 64, 64, seed=2000, device='cuda', n_hyp=128)
buf = (ctypes.c_ulonglong * 16)()
for rep in range(3):
    lib.posefit_debug_ransac_phases(buf, 1)
    pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], sample_idx=d['sample_idx'])
    lib.posefit_debug_ransac_phases(buf, 0)
v = np.array(list(buf)[:len(names)], dtype=np.float64)
print(f'cycles per object (thread 0), {b} objects: total {v.sum() / b:.0f}')
for n, x in zip(names, v):
    print(f'  {n:34s} {x / b:9.0f}  {100 * x / v.sum():5.1f} %')
