#!/usr/bin/env python
"""A/B of launch knobs on the config-5 shard, alternating the variants in one process (tooling, like tests/).  Default: the
ticketed (POSEFIT_DYNAMIC=1) against the fixed assignment of work.  A variant is a comma-separated list of POSEFIT_*
assignments without the prefix (`DYNAMIC=0`, `BWD_CTAS_PER_SM=3,BWD_MINB=3`); `base` = the defaults.  Prints the forward /
backward interval per variant, median over --steps steps per round."""
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
    ap.add_argument('--objects', type=int, default=125000)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--rounds', type=int, default=4)
    ap.add_argument('variants', nargs='*', default=['DYNAMIC=1', 'DYNAMIC=0'])
    a = ap.parse_args()
    dev = torch.device('cuda')
    kinv = pf.default_kinv(dev)
    n = a.objects
    c = pf.synth.make_objects(n, 64, 64, seed=5000, device=dev)
    g = (torch.randn(n, device=dev), torch.randn(n, 9, device=dev), torch.randn(n, 3, device=dev))
    gn = torch.empty_like(c['noc'])
    for rnd in range(a.rounds):
        for var in a.variants:
            for k in [k for k in os.environ if k.startswith('POSEFIT_') and k != 'POSEFIT_LIB']:
                del os.environ[k]
            if var != 'base':
                for kv in var.split(','):
                    k, v = kv.split('=')
                    os.environ['POSEFIT_' + k] = v
            pf._lib.reload_knobs()
            fw, bw = [], []
            for s in range(a.steps + 3):
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                raw = pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv)
                e1.record()
                pf.pose_fit_backward_raw(c['noc'], c['depth'], c['mask'], None, c['bbox_xy0'], kinv, raw.ctx, raw.status, *g, out=gn)
                e2.record()
                torch.cuda.synchronize()
                if s >= 3:
                    fw.append(e0.elapsed_time(e1))
                    bw.append(e1.elapsed_time(e2))
            fw.sort(); bw.sort()
            print(f'round {rnd} {var:28s}: forward {fw[len(fw) // 2]:.4f} ms, backward {bw[len(bw) // 2]:.4f} ms, '
                  f'step {fw[len(fw) // 2] + bw[len(bw) // 2]:.4f} ms', flush=True)


if __name__ == '__main__':
    main()
