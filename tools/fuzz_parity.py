#!/usr/bin/env python
"""Randomised parity run of the CUDA path against the NumPy oracle (test infrastructure, like tests/).

Every case draws a crop shape, batch size, mask density, hypothesis count, sample size, x0 alignment and
an optional unaligned storage offset, then checks plain fit, RANSAC fit (inlier masks bit-exact, winner,
statuses, poses within the north-star tolerance) and that both backward passes finish with finite
gradients that vanish outside the fit.  The case is printed BEFORE it runs, so a hang (run this under
`timeout`) names its shape.  usage: python tools/fuzz_parity.py [--cases N] [--seed S] [--seconds T]"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import posefit_oracle as po  # noqa: E402

pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')

ROT_TOL_DEG, REL_TOL = 1e-3, 1e-5
ties = [0]


def rot_err_deg(ra, rb):
    return float(np.degrees(np.linalg.norm(ra - rb) / np.sqrt(2)))


def compare(raw, ora, ransac, tag, sample_idx=None):
    pose = raw.pose.cpu().numpy()
    status = raw.status.cpu().numpy()
    n_valid = raw.n_valid.cpu().numpy()
    inl = raw.inlier_mask.cpu().numpy() if raw.inlier_mask is not None else None
    worst = 0.0
    for i, o in enumerate(ora):
        assert n_valid[i] == o['n_valid'], (tag, 'n_valid', i)
        if ransac and 0 < o['n_valid'] < 3:
            # one or two correspondences: EVERY hypothesis has a rank <= 1 covariance, the rotation (and with it the
            # inlier ratio gate) is LAPACK's arbitrary choice -- nothing to compare beyond the count
            ties[0] += 1
            continue
        assert status[i] == o['status'], (tag, 'status', i, int(status[i]), o['status'])
        tie = False
        if ransac and 'residuals' in o:
            # two hypotheses whose residuals agree to the last bits (the same samples drawn in another order)
            # are ranked by rounding noise in ANY arithmetic: not a parity statement
            r = np.sort(np.asarray(o['residuals'])[np.isfinite(o['residuals'])])
            tie = r.size > 1 and (r[1] - r[0]) <= 1e-9 * r[0]
            # a hypothesis drawn from fewer than 3 distinct points has a rank <= 1 covariance: its rotation is whatever
            # LAPACK's SVD happens to return (any rotation about the one determined axis is optimal)
            for hyp in (o.get('winner', -1), int(raw.winner[i])):
                if hyp >= 0 and len(set(sample_idx[i, hyp].tolist())) < 3:
                    tie = True
            ties[0] += int(tie)
        if ransac and not tie and o['status'] in (0, 2) and o.get('margin', np.inf) >= 1e-6:
            assert np.array_equal(inl[i], o['inlier_mask']), (tag, 'inlier mask', i, o.get('margin'))
            if o['status'] == 0:
                assert int(raw.winner[i]) == o['winner'], (tag, 'winner', i)
        n_fit = len(o['inlier_idx']) if 'inlier_idx' in o else o['n_valid']
        if o['status'] != 0 or tie or n_fit < 12:
            continue                                   # tiny clouds: singular covariances, statuses only
        e_r = rot_err_deg(pose[i, 1:10].reshape(3, 3), o['R'])
        e_t = float(np.linalg.norm(pose[i, 10:13] - o['t']) / max(np.linalg.norm(o['t']), 1e-30))
        e_s = abs(pose[i, 0] - o['s']) / abs(o['s'])
        assert e_r <= ROT_TOL_DEG and e_t <= REL_TOL and e_s <= REL_TOL, (tag, 'pose', i, e_r, e_t, e_s, o['n_valid'])
        worst = max(worst, e_r / ROT_TOL_DEG, e_t / REL_TOL, e_s / REL_TOL)
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=200)
    ap.add_argument('--seed', type=int, default=1)
    ap.add_argument('--seconds', type=float, default=90.0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0 = time.time()
    worst = 0.0
    done = failed = 0
    for case in range(a.cases):
        if time.time() - t0 > a.seconds:
            break
        kind = rng.integers(0, 4)
        if kind == 0:      # small / odd shapes
            h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        elif kind == 1:    # vector-path shapes (W % 4 == 0)
            h, w = int(rng.integers(1, 33)) * 4, int(rng.integers(1, 33)) * 4
        elif kind == 2:    # anything up to beyond the shared-memory staging of the RANSAC kernel
            h, w = int(rng.integers(20, 150)), int(rng.integers(20, 180))
        else:              # the benchmark shapes and their neighbours
            h, w = [(64, 64), (112, 112), (64, 60), (63, 64), (128, 128), (32, 128)][int(rng.integers(0, 6))]
        b = int(rng.integers(1, 7)) if h * w > 6000 else int(rng.integers(1, 24))
        fill = float(rng.choice([0.0, 0.02, 0.3, 0.7, 1.0], p=[0.05, 0.1, 0.2, 0.45, 0.2]))
        zero = float(rng.choice([0.0, 0.02, 0.5]))
        border = int(rng.integers(0, 3)) if min(h, w) > 6 else 0
        align = 4 if (w % 4 == 0 and rng.random() < 0.7) else 1
        n_hyp = int(rng.choice([1, 7, 32, 100, 128, 129, 200]))
        n_samp = int(rng.choice([3, 10, 10, 16]))
        outl = float(rng.choice([0.0, 0.1, 0.4]))
        off = int(rng.choice([0, 0, 1, 3]))          # element offset of the views into their storage
        print(f'case {case}: h={h} w={w} b={b} fill={fill} zero={zero} border={border} align={align} n_hyp={n_hyp} '
              f'n_samp={n_samp} outl={outl} off={off}', flush=True)
        d = pf.synth.make_objects(b, h, w, seed=int(rng.integers(1 << 30)), n_hyp=n_hyp, n_samp=n_samp, align_x0=align,
                                  border=border, mask_fill=fill, zero_depth_frac=zero, outlier_frac=outl)

        def dev(x):
            flat = torch.empty(x.numel() + off, dtype=x.dtype, device='cuda')
            v = flat[off:].view(x.shape)
            v.copy_(x)
            return v
        noc, depth, mask = dev(d['noc']), dev(d['depth']), dev(d['mask'])
        xy0, idx = d['bbox_xy0'].cuda(), d['sample_idx'].cuda()
        np_in = (d['noc'].numpy(), d['depth'].numpy(), d['mask'].numpy(), d['bbox_xy0'].numpy())

        try:
            raw = pf.pose_fit_raw(noc, depth, mask, xy0)
            torch.cuda.synchronize()
            worst = max(worst, compare(raw, po.batch_pose(*np_in), False, 'plain'))
            rr = pf.pose_fit_raw(noc, depth, mask, xy0, sample_idx=idx)
            torch.cuda.synchronize()
            worst = max(worst, compare(rr, po.batch_pose(*np_in, sample_idx=d['sample_idx'].numpy()), True, 'ransac',
                                            d['sample_idx'].numpy()))
            for use_idx in (None, idx):
                nq = noc.clone().requires_grad_(True)
                dq = depth.clone().requires_grad_(True)
                out = pf.pose_fit(nq, dq, mask, xy0, sample_idx=use_idx)
                (out[0].sum() + out[1].sum() + out[2].sum()).backward()
                torch.cuda.synchronize()
                assert torch.isfinite(nq.grad).all() and torch.isfinite(dq.grad).all(), 'non-finite gradient'
                used = (mask != 0) & (depth > 0)
                if use_idx is not None:
                    used = used & (out[3] != 0)
                assert (nq.grad * (~used)[:, None]).abs().max().item() == 0.0, 'gradient outside the fit'
                assert (dq.grad * ~used).abs().max().item() == 0.0, 'depth gradient outside the fit'
        except AssertionError as e:
            failed += 1
            print('  FAILED:', e, flush=True)
        done += 1
    print(f'fuzz {"ok" if failed == 0 else "FAILED (%d cases)" % failed}: {done} cases in {time.time() - t0:.1f} s, worst error / tolerance = {worst:.3g}, {ties[0]} tied objects skipped')
    sys.exit(1 if failed else 0)


if __name__ == '__main__':
    main()
