#!/usr/bin/env python
"""fit_ransac_crop_kernel against fit_ransac_kernel (POSEFIT_RANSAC_SCREEN=0) on many seeded batches; every
disagreement is printed with the oracle's residuals of the two winners (test infrastructure, like tests/).
usage: python tools/crop_vs_general.py [--batches N] [--threads 128,160,...]"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import posefit_oracle as po  # noqa: E402

pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')


def knob(**kw):
    for k in [k for k in os.environ if k.startswith('POSEFIT_')]:
        del os.environ[k]
    for k, v in kw.items():
        os.environ['POSEFIT_' + k] = str(v)
    pf._lib.reload_knobs()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, default=40)
    ap.add_argument('--threads', default='128')
    a = ap.parse_args()
    rng = np.random.default_rng(5)
    regimes = [dict(), dict(outlier_frac=0.0), dict(outlier_frac=0.0, noc_noise=0.0), dict(outlier_frac=0.4),
               dict(mask_fill=0.3), dict(mask_fill=0.1, border=0), dict(zero_depth_frac=0.5)]
    bad = total = 0
    for it in range(a.batches):
        reg = regimes[it % len(regimes)]
        h, w = [(64, 64), (32, 32), (48, 64), (40, 52), (112, 112), (64, 128)][int(rng.integers(0, 6))]
        b = int(rng.integers(1, 700)) if h * w <= 4096 else int(rng.integers(1, 40))
        n_hyp = int(rng.choice([1, 7, 32, 100, 128, 200]))
        n_samp = int(rng.choice([3, 10, 10, 16]))
        seed = int(rng.integers(1 << 30))
        d = pf.synth.make_objects(b, h, w, seed=seed, n_hyp=n_hyp, n_samp=n_samp, **reg)
        t = {k: d[k].cuda() for k in ('noc', 'depth', 'mask', 'bbox_xy0', 'sample_idx')}
        knob(RANSAC_SCREEN=0)
        ref = pf.pose_fit_raw(t['noc'], t['depth'], t['mask'], t['bbox_xy0'], sample_idx=t['sample_idx'])
        torch.cuda.synchronize()
        for nt in a.threads.split(','):
            knob(RANSAC_THREADS=nt)
            out = pf.pose_fit_raw(t['noc'], t['depth'], t['mask'], t['bbox_xy0'], sample_idx=t['sample_idx'])
            torch.cuda.synchronize()
            total += b
            dw = (out.winner != ref.winner).nonzero().flatten().tolist()
            dm = (out.inlier_mask != ref.inlier_mask).flatten(1).any(1).nonzero().flatten().tolist()
            ds = (out.status != ref.status).nonzero().flatten().tolist()
            dp = float((out.pose[:, :15] - ref.pose[:, :15]).abs().max())
            if dw or dm or ds or dp > 1e-9:
                print(f'batch {it}: {h}x{w} b={b} n_hyp={n_hyp} n_samp={n_samp} {reg} nt={nt}: winners differ {dw[:5]} '
                      f'masks differ {dm[:5]} status differ {ds[:5]} dpose {dp:.2e}', flush=True)
                for i in (dw + dm)[:3]:
                    o = po.batch_pose(d['noc'][i:i + 1].numpy(), d['depth'][i:i + 1].numpy(), d['mask'][i:i + 1].numpy(),
                                      d['bbox_xy0'][i:i + 1].numpy(), sample_idx=d['sample_idx'][i:i + 1].numpy())[0]
                    res = np.asarray(o.get('residuals', []))
                    wa, wb = int(out.winner[i]), int(ref.winner[i])
                    print(f'   object {i}: crop winner {wa} general {wb} oracle {o.get("winner")} n_valid {o["n_valid"]} '
                          f'status {int(out.status[i])}/{int(ref.status[i])}/{o["status"]} '
                          f'oracle residuals {res[wa] if 0 <= wa < res.size else None!r} '
                          f'{res[wb] if 0 <= wb < res.size else None!r} margin {o.get("margin")} '
                          f'samples {d["sample_idx"][i, wa].tolist() if wa >= 0 else None} '
                          f'{d["sample_idx"][i, wb].tolist() if wb >= 0 else None} '
                          f'mask diff px {int((out.inlier_mask[i] != ref.inlier_mask[i]).sum())}')
                bad += 1
    print(f'crop vs general: {total} object runs, {bad} batches with a disagreement')


if __name__ == '__main__':
    main()
