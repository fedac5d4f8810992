#!/usr/bin/env python
"""Generate tests/golden/grad_fd.npz: gradients of the pose fit obtained by CENTRAL FINITE
DIFFERENCES OF THE UNMODIFIED REFERENCE FUNCTIONS from /root/reference.

TEST INFRASTRUCTURE.  Run in the build container only (`python oracle/gen_golden_grad.py`).

The reference has no backward pass (it detaches the NOC patch, Detection/tracker/postprocess.py:151),
so there is no reference gradient to copy.  What the reference does define is the FUNCTION
    (noc crop, depth crop) -> (s, R, t)
    = estimateSimilarityUmeyama(nocs[rows, cols] - 0.5, backproject(depth_pad, K, mask_pad))
(pose_estimation.py:256-290, :323, pose_utils.py:16-61).  For a linear probe
    L = g_s s + <G_R, R> + <g_t, t>
its derivative with respect to every NOC value and every depth value of the crop is evaluated here
with central differences (float64, step 1e-6: truncation ~1e-12, round-off ~1e-10 relative) calling
the real `PoseEst.pose_estimation.backproject` and `PoseEst.pose_utils.estimateSimilarityUmeyama`
for every perturbed input.  These vectors pin the gradient oracle (oracle/grad_oracle.py), the
host-compiled adjoint (csrc/posefit_math.h) and the CUDA backward kernel.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import ref_import  # noqa: E402
from oracle import posefit_oracle as po  # noqa: E402  (only for the fixed MOTFront intrinsics matrix)

GOLD = os.path.join(ROOT, 'tests', 'golden')
STEP = 1e-6


def _hom(p):
    return np.transpose(np.hstack([p, np.ones([p.shape[0], 1])]))


def make_case(rng, h, w, x0, y0, fill):
    """A small MOTFront-shaped instance: NOC in [0,1], depth consistent with a random similarity."""
    k = po.motfront_intrinsics()
    kinv = np.linalg.inv(k)
    noc = rng.uniform(0.05, 0.95, size=(h, w, 3)).astype(np.float32)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    rot = np.array([[1 - 2 * (c * c + d * d), 2 * (b * c - d * a), 2 * (b * d + c * a)],
                    [2 * (b * c + d * a), 1 - 2 * (b * b + d * d), 2 * (c * d - b * a)],
                    [2 * (b * d - c * a), 2 * (c * d + b * a), 1 - 2 * (b * b + c * c)]])
    s, t = rng.uniform(0.6, 2.0), np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), -rng.uniform(2.5, 4.0)])
    pts = s * (noc.reshape(-1, 3).astype(np.float64) - 0.5) @ rot.T + t          # camera space, z < 0 (:40-41 flips)
    depth = (-pts[:, 2]).reshape(h, w) + rng.normal(scale=0.02, size=(h, w))    # a noisy depth map, NOT exactly consistent
    depth = depth.astype(np.float32)
    mask = (rng.uniform(size=(h, w)) < fill)
    mask[0, :] = False
    depth[rng.uniform(size=(h, w)) < 0.05] = 0.0                                 # holes: depth > 0 test (:23-25)
    del kinv
    return noc, depth, mask.astype(np.uint8), np.array([x0, y0], dtype=np.int32)


def probe(pu, pe, noc, depth, mask, xy0, g_s, g_r, g_t, sample_idx=None):
    """L(noc, depth) through the real reference functions; noc [h,w,3], depth [h,w] float64.
    sample_idx [n_hyp,10]: go through estimateSimilarityTransform (RANSAC, pose_utils.py:86-117) with the
    indices replayed instead of the plain fit."""
    h, w = depth.shape
    x0, y0 = int(xy0[0]), int(xy0[1])
    depth_pad = np.zeros((po.FRAME_H, po.FRAME_W))                               # pose_estimation.py:260-262
    depth_pad[y0:y0 + h, x0:x0 + w] = depth
    mask_pad = np.zeros((po.FRAME_H, po.FRAME_W), dtype=bool)
    mask_pad[y0:y0 + h, x0:x0 + w] = mask != 0
    nocs_pad = np.zeros((po.FRAME_H, po.FRAME_W, 3))                             # :265-267
    nocs_pad[y0:y0 + h, x0:x0 + w, :] = noc
    pts, idxs = pe.backproject(depth_pad, po.motfront_intrinsics(), mask_pad)   # :290
    noc_pts = nocs_pad[idxs[0], idxs[1], :] - 0.5                                # :323
    if sample_idx is None:
        scales, rotation, translation, _ = pu.estimateSimilarityUmeyama(_hom(noc_pts), _hom(pts))
    else:
        with ref_import.replay_randint(sample_idx, pu):
            scales, rotation, translation, _ = pu.estimateSimilarityTransform(noc_pts, pts)
    s, rot, t = scales[0], rotation.T, translation                               # Rotation is R^T (pose_utils.py:44)
    return g_s * s + float((g_r * rot).sum()) + float((g_t * t).sum()), (s, rot, t), int(pts.shape[0])


def finite_differences(pu, pe, noc32, depth32, mask, xy0, g_s, g_r, g_t, sample_idx=None):
    noc = noc32.astype(np.float64)
    depth = depth32.astype(np.float64)
    h, w = depth.shape
    _, fwd, n_valid = probe(pu, pe, noc, depth, mask, xy0, g_s, g_r, g_t, sample_idx)
    valid = (mask != 0) & (depth > 0)
    g_noc = np.zeros((h, w, 3))
    g_depth = np.zeros((h, w))
    for i in range(h):
        for j in range(w):
            if not valid[i, j]:
                continue                              # an invalid pixel never reaches the fit: exact zero
            for c in range(3):
                p, m = noc.copy(), noc.copy()
                p[i, j, c] += STEP
                m[i, j, c] -= STEP
                g_noc[i, j, c] = (probe(pu, pe, p, depth, mask, xy0, g_s, g_r, g_t, sample_idx)[0] -
                                  probe(pu, pe, m, depth, mask, xy0, g_s, g_r, g_t, sample_idx)[0]) / (2 * STEP)
            p, m = depth.copy(), depth.copy()
            p[i, j] += STEP
            m[i, j] -= STEP
            g_depth[i, j] = (probe(pu, pe, noc, p, mask, xy0, g_s, g_r, g_t, sample_idx)[0] -
                             probe(pu, pe, noc, m, mask, xy0, g_s, g_r, g_t, sample_idx)[0]) / (2 * STEP)
    return g_noc, g_depth, fwd, n_valid


def main():
    pu, pe = ref_import.load_reference()
    rng = np.random.default_rng(20261018)
    out = {}
    cases = [(12, 16, 100, 60, 0.8), (9, 11, 37, 151, 0.7), (16, 8, 208, 20, 0.35), (6, 20, 3, 200, 0.9)]
    for k, (h, w, x0, y0, fill) in enumerate(cases):
        noc, depth, mask, xy0 = make_case(rng, h, w, x0, y0, fill)
        g_s = rng.normal()
        g_r = rng.normal(size=(3, 3))
        g_t = rng.normal(size=3)
        g_noc, g_depth, (s, rot, t), n_valid = finite_differences(pu, pe, noc, depth, mask, xy0, g_s, g_r, g_t)
        out[f'noc_{k}'] = np.ascontiguousarray(np.transpose(noc, (2, 0, 1)))           # planar [3,h,w] f32
        out[f'depth_{k}'], out[f'mask_{k}'], out[f'xy0_{k}'] = depth, mask, xy0
        out[f'g_s_{k}'], out[f'g_R_{k}'], out[f'g_t_{k}'] = np.float64(g_s), g_r, g_t
        out[f'grad_noc_{k}'] = np.ascontiguousarray(np.transpose(g_noc, (2, 0, 1)))    # [3,h,w] f64
        out[f'grad_depth_{k}'] = g_depth
        out[f's_{k}'], out[f'R_{k}'], out[f't_{k}'], out[f'n_valid_{k}'] = np.float64(s), rot, t, np.int32(n_valid)
        print(f'case {k}: {h}x{w} n_valid {n_valid} |grad_noc| {np.abs(g_noc).max():.3e} |grad_depth| {np.abs(g_depth).max():.3e}')
    # RANSAC path: gradient through the refit on the winner's inliers (winner and inlier set are locally
    # constant, so the finite differences see the same piecewise-smooth function the backward kernel differentiates)
    k = len(cases)
    noc, depth, mask, xy0 = make_case(rng, 14, 16, 64, 90, 0.85)
    valid = (mask != 0) & (depth > 0)
    bad = valid & (rng.uniform(size=depth.shape) < 0.15)
    depth[bad] += rng.uniform(8.0, 20.0, size=int(bad.sum())).astype(np.float32)     # gross outliers, beyond PassT
    n_valid = int(valid.sum())
    sample_idx = rng.integers(0, n_valid, size=(12, 10)).astype(np.int32)
    g_s, g_r, g_t = rng.normal(), rng.normal(size=(3, 3)), rng.normal(size=3)
    g_noc, g_depth, (s, rot, t), n_valid2 = finite_differences(pu, pe, noc, depth, mask, xy0, g_s, g_r, g_t, sample_idx)
    assert n_valid2 == n_valid
    out[f'noc_{k}'] = np.ascontiguousarray(np.transpose(noc, (2, 0, 1)))
    out[f'depth_{k}'], out[f'mask_{k}'], out[f'xy0_{k}'] = depth, mask, xy0
    out[f'g_s_{k}'], out[f'g_R_{k}'], out[f'g_t_{k}'] = np.float64(g_s), g_r, g_t
    out[f'grad_noc_{k}'] = np.ascontiguousarray(np.transpose(g_noc, (2, 0, 1)))
    out[f'grad_depth_{k}'] = g_depth
    out[f's_{k}'], out[f'R_{k}'], out[f't_{k}'], out[f'n_valid_{k}'] = np.float64(s), rot, t, np.int32(n_valid)
    out[f'sample_idx_{k}'] = sample_idx
    n_zero = int(((np.abs(g_noc).sum(-1) == 0) & valid).sum())
    print(f'case {k} (RANSAC): n_valid {n_valid}, {n_zero} valid pixels with zero gradient (outliers), '
          f'|grad_noc| {np.abs(g_noc).max():.3e}')
    out['ransac_case'] = np.int32(k)
    out['n_cases'] = np.int32(len(cases))
    out['step'] = np.float64(STEP)
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, 'grad_fd.npz'), **out)
    print('wrote', os.path.join(GOLD, 'grad_fd.npz'))


if __name__ == '__main__':
    main()
