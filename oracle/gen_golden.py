#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE.  Run in the build container only (`python oracle/gen_golden.py`);
the fixtures it writes are committed, the reference tree is not.  Every output array in
the fixtures comes from a call into the real `PoseEst.pose_utils` /
`PoseEst.pose_estimation` functions; `np.random.randint` is patched to replay the stored
sample indices (pose_utils.py:73 draws from the global unseeded RNG otherwise).
"""
from __future__ import annotations

import importlib
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def _hom(p):
    return np.transpose(np.hstack([p, np.ones([p.shape[0], 1])]))


def _rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def umeyama_cases(pu, rng):
    cases = []

    def add(name, src, dst):
        s, r, t, tf = pu.estimateSimilarityUmeyama(_hom(src), _hom(dst))
        cases.append((name, src, dst, s, r, t, tf))

    for i, n in enumerate((4, 10, 10, 57, 1000, 2900)):
        src = rng.uniform(-0.5, 0.5, size=(n, 3))
        rot, sc, tr = _rand_rot(rng), rng.uniform(0.5, 2.5), rng.uniform(-3, 3, size=3)
        dst = sc * src @ rot.T + tr
        add(f'exact_{i}', src, dst)
        add(f'noisy_{i}', src, dst + rng.normal(scale=0.02, size=dst.shape))
    # reflection branch pose_utils.py:39-42
    src = rng.uniform(-0.5, 0.5, size=(200, 3))
    dst = 1.7 * src @ _rand_rot(rng).T * np.array([1.0, 1.0, -1.0]) + np.array([0.1, -0.4, -3.0])
    add('reflection', src, dst + rng.normal(scale=0.01, size=dst.shape))
    # near planar source and target
    src = rng.uniform(-0.5, 0.5, size=(300, 3)) * np.array([1.0, 1.0, 1e-3])
    add('planar', src, 1.2 * src @ _rand_rot(rng).T + np.array([0.3, 0.2, -2.5]) + rng.normal(scale=1e-3, size=src.shape))
    # zero variance: all points equal -> scale 1 (pose_utils.py:47-50)
    src = np.tile(rng.uniform(-0.5, 0.5, size=(1, 3)), (10, 1))
    dst = np.tile(rng.uniform(-3, 3, size=(1, 3)), (10, 1))
    add('roundoff_variance', src, dst)       # mean of 10 equal values is not exact -> garbage scale, chaotic
    add('single_point', src[:1], dst[:1])
    # exactly zero variance (dyadic values, 8 points: every sum is exact) -> scale 1, rotation I
    add('zero_variance', np.tile(np.array([[0.25, -0.5, 0.125]]), (8, 1)), np.tile(np.array([[1.5, -0.75, -3.0]]), (8, 1)))
    # repeated sample indices, as RANSAC with replacement produces
    base_s = rng.uniform(-0.5, 0.5, size=(4, 3))
    base_d = 0.9 * base_s @ _rand_rot(rng).T + np.array([0.0, 0.5, -4.0]) + rng.normal(scale=0.05, size=(4, 3))
    pick = np.array([0, 1, 1, 2, 3, 3, 3, 0, 2, 1])
    add('repeated', base_s[pick], base_d[pick])
    out = {'names': np.array([c[0] for c in cases])}
    for k, c in enumerate(cases):
        out[f'src_{k}'], out[f'dst_{k}'] = c[1], c[2]
        out[f'scales_{k}'], out[f'rotation_{k}'], out[f'translation_{k}'], out[f'transform_{k}'] = c[3:]
    return out


def evaluate_cases(pu, rng):
    out = {}
    k = 0
    for n in (1, 7, 500):
        src = rng.uniform(-0.5, 0.5, size=(n, 3))
        dst = rng.uniform(-4, 4, size=(n, 3))
        tf = np.identity(4)
        tf[:3, :3] = rng.uniform(0.5, 2) * _rand_rot(rng)
        tf[:3, 3] = rng.uniform(-1, 1, size=3)
        r = np.linalg.norm((dst.T - (tf[:3, :3] @ src.T + tf[:3, 3:4])), axis=0)
        for pass_t in (np.median(r), r.max() + 1.0, r.min() * 0.5):
            res, ratio, idx = pu.evaluateModel(tf, _hom(src), _hom(dst), pass_t)
            out[f'src_{k}'], out[f'dst_{k}'], out[f'transform_{k}'] = src, dst, tf
            out[f'pass_{k}'], out[f'residual_{k}'], out[f'ratio_{k}'], out[f'idx_{k}'] = pass_t, res, ratio, idx
            k += 1
    out['count'] = np.array(k)
    return out


def _make_cloud(rng, n, outlier_frac, rot=None, out_lo=8.0, out_hi=20.0, noise=0.01):
    src = rng.uniform(-0.5, 0.5, size=(n, 3))
    rot = _rand_rot(rng) if rot is None else rot
    dst = rng.uniform(0.6, 2.2) * src @ rot.T + np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), -rng.uniform(2.5, 4.5)])
    dst += rng.normal(scale=noise, size=dst.shape)
    bad = rng.uniform(size=n) < outlier_frac
    dst[bad, 2] -= rng.uniform(out_lo, out_hi, size=bad.sum())
    return src, dst


def ransac_cases(pu, rng):
    """estimateSimilarityTransform end to end (pose_utils.py:86-117) with replayed indices."""
    out = {}
    specs = [
        ('generic_a', 400, 0.10, None, 24), ('generic_b', 900, 0.25, None, 32), ('generic_c', 2900, 0.10, None, 100),
        ('clean', 300, 0.0, None, 16),
        ('identity_rot_early_stop', 500, 0.0, np.identity(3), 12),      # F3: only R~I can reach StopT
        ('mostly_outliers', 300, 0.97, None, 16),
        ('tiny', 12, 0.2, None, 8),
    ]
    k = 0
    for name, n, frac, rot, n_hyp in specs:
        src, dst = _make_cloud(rng, n, frac, rot, noise=0.0 if name.startswith('identity') else 0.01)
        if name == 'mostly_outliers':        # unrelated clouds of equal norm -> PassT ~ 1, few inliers -> gate
            src, dst = rng.uniform(-5, 5, size=(n, 3)), rng.uniform(-5, 5, size=(n, 3))
        idx = rng.integers(0, n, size=(n_hyp, 10))
        if name == 'generic_a':
            idx[5] = idx[2]                 # exact tie -> first wins (strict <, pose_utils.py:76)
        with ref_import.replay_randint(idx, pu) as calls, redirect_stdout(io.StringIO()):
            s, r, t, tf = pu.estimateSimilarityTransform(src, dst, verbose=False)
        out[f'name_{k}'] = np.array(name)
        out[f'src_{k}'], out[f'dst_{k}'], out[f'idx_{k}'] = src, dst, idx
        out[f'calls_{k}'] = np.array(calls['n'])
        out[f'ok_{k}'] = np.array(s is not None)
        if s is not None:
            out[f'scales_{k}'], out[f'rotation_{k}'], out[f'translation_{k}'], out[f'transform_{k}'] = s, r, t, tf
        k += 1
    out['count'] = np.array(k)
    return out


def ransac_internal_cases(pu, rng):
    """getRANSACInliers alone (pose_utils.py:63-83): selected inlier sets and ratios."""
    out = {}
    k = 0
    for n, frac, n_hyp in ((200, 0.15, 20), (1500, 0.3, 40), (64, 0.5, 10)):
        src, dst = _make_cloud(rng, n, frac)
        idx = rng.integers(0, n, size=(n_hyp, 10))
        t_norm = np.mean(np.linalg.norm(dst, axis=1))
        s_norm = np.mean(np.linalg.norm(src, axis=1))
        pass_t = max(t_norm / s_norm, s_norm / t_norm)
        with ref_import.replay_randint(idx):
            s_in, d_in, ratio = pu.getRANSACInliers(_hom(src), _hom(dst), MaxIterations=n_hyp,
                                                    PassThreshold=pass_t, StopThreshold=pass_t / 100)
        out[f'src_{k}'], out[f'dst_{k}'], out[f'idx_{k}'] = src, dst, idx
        out[f'pass_{k}'] = np.array(pass_t)
        out[f'src_in_{k}'], out[f'dst_in_{k}'], out[f'ratio_{k}'] = s_in[:3].T, d_in[:3].T, np.array(ratio)
        k += 1
    out['count'] = np.array(k)
    return out


def frame_cases(pu, pe):
    """BASELINE config 1: 8 objects with 64x64 crops in one 240x320 frame, through the real
    backproject (pose_estimation.py:16-43), the run_pose padding/gather (:256-267, :323) restated
    around it, and the real fits.  Inputs are stored in the crop layout the CUDA path reads."""
    import torch
    synth = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.synth')
    out = {}
    for tag, (h, w, b, n_hyp, seed) in {'c1': (64, 64, 8, 32, 101), 'small': (24, 32, 6, 16, 102),
                                        'odd': (19, 27, 4, 12, 103)}.items():
        d = synth.make_objects(b, h, w, seed=seed, n_hyp=n_hyp, align_x0=4 if tag != 'odd' else 1,
                               outlier_range=(25.0, 40.0) if tag != 'small' else (8.0, 20.0))
        noc, depth, mask = d['noc'].numpy(), d['depth'].numpy(), d['mask'].numpy()
        xy0, idx = d['bbox_xy0'].numpy(), d['sample_idx'].numpy()
        if tag == 'small':
            mask[1] = 0                              # empty object -> 6xNone (pose_estimation.py:361-362)
        k_mat = synth.motfront_intrinsics().numpy()
        out[f'{tag}_noc'], out[f'{tag}_depth'], out[f'{tag}_mask'] = noc, depth, mask
        out[f'{tag}_bbox_xy0'], out[f'{tag}_sample_idx'], out[f'{tag}_K'] = xy0, idx, k_mat
        for i in range(b):
            x0, y0 = int(xy0[i, 0]), int(xy0[i, 1])
            depth_pad = np.zeros((240, 320))
            depth_pad[y0:y0 + h, x0:x0 + w] = depth[i]
            noc_pad = np.zeros((240, 320, 3))
            noc_pad[y0:y0 + h, x0:x0 + w, :] = np.transpose(noc[i], (1, 2, 0))
            mask_pad = np.zeros((240, 320), dtype=bool)
            mask_pad[y0:y0 + h, x0:x0 + w] = mask[i] != 0
            pts, idxs = pe.backproject(depth_pad, k_mat, mask_pad)
            noc_pts = noc_pad[idxs[0], idxs[1], :] - 0.5
            out[f'{tag}_{i}_n_valid'] = np.array(pts.shape[0])
            if tag != 'c1':                          # keep the committed fixture small
                out[f'{tag}_{i}_pts'], out[f'{tag}_{i}_noc_pts'] = pts, noc_pts
                out[f'{tag}_{i}_rows'], out[f'{tag}_{i}_cols'] = idxs[0].astype(np.int16), idxs[1].astype(np.int16)
            if pts.shape[0] == 0:
                out[f'{tag}_{i}_status'] = np.array(1)
                continue
            s, r, t, tf = pu.estimateSimilarityUmeyama(_hom(noc_pts), _hom(pts))
            out[f'{tag}_{i}_fit_scales'], out[f'{tag}_{i}_fit_rotation'] = s, r
            out[f'{tag}_{i}_fit_translation'] = t
            n = pts.shape[0]
            ii = np.minimum(idx[i], n - 1)
            with ref_import.replay_randint(ii, pu) as calls, redirect_stdout(io.StringIO()):
                s, r, t, tf = pu.estimateSimilarityTransform(noc_pts, pts, verbose=False)
            out[f'{tag}_{i}_status'] = np.array(0 if s is not None else 2)
            # inlier mask of the winning hypothesis, in crop coordinates
            src_in, dst_in, ratio = calls['ransac_out']
            both = np.vstack([_hom(noc_pts), _hom(pts)])
            keep = ref_import.inlier_indices_from_subset(both, np.vstack([src_in, dst_in]))
            im = np.zeros((h, w), dtype=np.uint8)
            im[idxs[0][keep] - y0, idxs[1][keep] - x0] = 1
            out[f'{tag}_{i}_inlier_mask'] = np.packbits(im)
            out[f'{tag}_{i}_ratio'] = np.array(ratio)
            out[f'{tag}_{i}_pass_t'] = np.array(calls['pass_t'])
            if s is not None:
                out[f'{tag}_{i}_ransac_scales'], out[f'{tag}_{i}_ransac_rotation'] = s, r
                out[f'{tag}_{i}_ransac_translation'] = t
                if tag != 'c1':
                    cam = pe.transform_pc(s, r, t, noc_pts)
                    out[f'{tag}_{i}_transformed_pc'] = cam
                    campose = np.identity(4)
                    campose[:3, :3] = _rand_rot(np.random.default_rng(seed + i))
                    campose[:3, 3] = [0.5, -1.0, 2.0]
                    out[f'{tag}_{i}_campose'] = campose
                    out[f'{tag}_{i}_world_pc'] = pe.cam2world(cam, campose)
    return out


def clip_cases(pe, rng):
    """clean_depth (pose_estimation.py:107-134) on random clouds / boxes / camera poses."""
    out = {}
    k = 0
    for n, frac in ((300, 0.6), (1200, 0.3), (50, 0.05), (5, 1.0)):
        pts = rng.normal(size=(n, 3)) * np.array([0.5, 0.4, 0.6]) + np.array([0.2, -0.1, -3.5])
        campose = np.identity(4)
        campose[:3, :3] = _rand_rot(rng)
        campose[:3, 3] = rng.normal(size=3)
        world = pe.cam2world(pts, campose)
        c, e = np.median(world, axis=0), world.std(axis=0) * (0.3 + 2.0 * frac)
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
        box = c + corners * e
        box = box[rng.permutation(8)]
        new_depth, used = pe.clean_depth(pts, box, campose)
        out[f'pts_{k}'], out[f'box_{k}'], out[f'campose_{k}'] = pts, box, campose
        out[f'used_{k}'] = np.asarray(used, dtype=np.int64)
        out[f'new_depth_{k}'] = np.asarray(new_depth, dtype=np.float64).reshape(-1, 3)
        k += 1
    out['count'] = np.array(k)
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    pu, pe = ref_import.load_reference()
    rng = np.random.default_rng(20261018)
    np.savez_compressed(os.path.join(GOLD, 'umeyama.npz'), **umeyama_cases(pu, rng))
    np.savez_compressed(os.path.join(GOLD, 'evaluate.npz'), **evaluate_cases(pu, rng))
    np.savez_compressed(os.path.join(GOLD, 'ransac.npz'), **ransac_cases(pu, rng))
    np.savez_compressed(os.path.join(GOLD, 'ransac_inliers.npz'), **ransac_internal_cases(pu, rng))
    np.savez_compressed(os.path.join(GOLD, 'frames.npz'), **frame_cases(pu, pe))
    np.savez_compressed(os.path.join(GOLD, 'clip.npz'), **clip_cases(pe, np.random.default_rng(77)))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == '__main__':
    main()
