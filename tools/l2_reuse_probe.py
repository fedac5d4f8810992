#!/usr/bin/env python
"""Does the backward pass find the forward pass's crops in L2?  (tooling, like tests/)

Times fit_backward (coefficients + streaming kernel) of a C4-shaped batch right after a forward pass over the SAME
crops and right after a forward pass over OTHER crops (eager launches, events between the calls).  Round-2 result
(profiles/r02_l2_reuse_probe.txt): 2 us apart even when all crops fit in L2 (39 MB) -- the backward kernel of a short
batch is bound by its per-unit latency chain, not by where its reads come from; an evict-last policy on the forward
pass's copies and an L2 prefetch ahead of the backward kernel's griddepcontrol.wait were built, measured (no change on
BASELINE config 4) and removed."""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--objects', type=int, default=384)
    ap.add_argument('--size', type=int, default=112)
    ap.add_argument('--sets', type=int, default=8)
    ap.add_argument('--reps', type=int, default=24)
    a = ap.parse_args()
    dev = torch.device('cuda')
    kinv = pf.default_kinv(dev)
    n, sz = a.objects, a.size
    sets = [pf.synth.make_objects(n, sz, sz, seed=4000 + i, device=dev) for i in range(a.sets)]
    g = (torch.randn(n, device=dev), torch.randn(n, 9, device=dev), torch.randn(n, 3, device=dev))
    mb = n * sz * sz * 17 / 2**20
    print(f'{n} objects {sz}x{sz}: {mb:.1f} MB of crops per batch, {n * sz * sz * 12 / 2**20:.1f} MB of gradients written')
    for keep in ('-',):
        ctxs = []
        for c in sets:
            raw = pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv)
            ctxs.append((raw.ctx.clone(), raw.status.clone()))
        torch.cuda.synchronize()
        res = {}
        for mode in ('same', 'other'):
            fw, bw = [], []
            for r in range(a.reps + 4):
                i = r % a.sets
                j = i if mode == 'same' else (i + a.sets // 2) % a.sets
                c, d = sets[i], sets[j]
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv)
                e1.record()
                pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, ctxs[j][0], ctxs[j][1], *g)
                e2.record()
                torch.cuda.synchronize()
                if r >= 4:
                    fw.append(e0.elapsed_time(e1) * 1e3)
                    bw.append(e1.elapsed_time(e2) * 1e3)
            fw.sort(); bw.sort()
            res[mode] = (fw[len(fw) // 2], bw[len(bw) // 2])
        print(f'backward after a forward over the same crops {res["same"][1]:6.1f} us, over other crops '
              f'{res["other"][1]:6.1f} us   (forward {res["same"][0]:.1f} / {res["other"][0]:.1f} us; eager launches, events '
              f'between the calls)', flush=True)


if __name__ == '__main__':
    main()
