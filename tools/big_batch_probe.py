#!/usr/bin/env python
"""BASELINE config 5 on ONE GPU (1 000 000 objects of 64x64): forward and backward timed separately, as one launch
sequence over the whole batch and as sub-batches of `--sub` objects over the same tensors (tooling, like tests/).

    python tools/big_batch_probe.py [--objects 1000000] [--sub 125000,250000] [--steps 4]
"""
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
    ap.add_argument('--objects', type=int, default=1000000)
    ap.add_argument('--sub', default='125000,250000')
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--b2b', type=int, default=0, help='also time this many steps queued back to back')
    ap.add_argument('--heat', type=float, default=0.0, help='seconds of continuous load before every measurement')
    a = ap.parse_args()
    dev = torch.device('cuda')
    kinv = pf.default_kinv(dev)
    n = a.objects
    d = pf.synth.make_objects(n, 64, 64, seed=5100, device=dev)
    g = (torch.randn(n, device=dev), torch.randn(n, 9, device=dev), torch.randn(n, 3, device=dev))
    grad = torch.empty_like(d['noc'])

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def sm_clock():
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
            return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
        except Exception:                                        # noqa: BLE001 -- tooling: the clock is a side note
            return -1, -1.0

    def heat(seconds):
        import time
        t0 = time.time()
        while time.time() - t0 < seconds:
            raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)
            pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status,
                                     *g, out=grad)
            torch.cuda.synchronize()

    def run(sub):
        fwd = bwd = 0.0
        if a.heat > 0:
            heat(a.heat)
        for it in range(a.steps + 1):
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            raws = []
            for s in range(0, n, sub):
                e = min(n, s + sub)
                raws.append(pf.pose_fit_raw(d['noc'][s:e], d['depth'][s:e], d['mask'][s:e], d['bbox_xy0'][s:e], kinv))
            e1.record()
            for i, s in enumerate(range(0, n, sub)):
                e = min(n, s + sub)
                pf.pose_fit_backward_raw(d['noc'][s:e], d['depth'][s:e], d['mask'][s:e], None, d['bbox_xy0'][s:e], kinv,
                                         raws[i].ctx, raws[i].status, g[0][s:e], g[1][s:e], g[2][s:e], out=grad[s:e])
            e2.record()
            torch.cuda.synchronize()
            if it > 0:
                fwd += e0.elapsed_time(e1)
                bwd += e1.elapsed_time(e2)
            del raws
        return fwd / a.steps, bwd / a.steps

    def back_to_back(k):
        import threading
        import time
        stop, clocks = threading.Event(), []

        def sample():
            while not stop.is_set():
                clocks.append(sm_clock())
                time.sleep(0.01)
        th = threading.Thread(target=sample)
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        th.start()
        t0 = time.time()
        e0.record()
        for _ in range(k):
            raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)
            pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status,
                                     *g, out=grad)
        e1.record()
        t_host = (time.time() - t0) * 1e3
        torch.cuda.synchronize()
        stop.set()
        th.join()
        mhz = sorted(c[0] for c in clocks)
        print(f'{k} steps back to back, no sync between: {e0.elapsed_time(e1) / k:.3f} ms per step; host enqueue of all steps '
              f'{t_host:.1f} ms; sm clock min / median / max {mhz[0]} / {mhz[len(mhz) // 2]} / {mhz[-1]} MHz, '
              f'power max {max(c[1] for c in clocks):.0f} W', flush=True)

    bf, bb = 17 * 4096 + 64, 29 * 4096 + 52
    if a.b2b > 0:
        back_to_back(2)
        back_to_back(a.b2b)
        back_to_back(a.b2b)
    for sub in [n] + [int(x) for x in a.sub.split(',') if x]:
        f, b = run(sub)
        mhz, watts = sm_clock()
        print(f'[sm {mhz} MHz, {watts:.0f} W right after] {n} objects in launches of {sub:8d}: fwd {f:7.3f} ms ({n * bf / f / 1e6:6.0f} GB/s)  bwd {b:7.3f} ms '
              f'({n * bb / b / 1e6:6.0f} GB/s)  step {f + b:7.3f} ms = {n / (f + b) / 1e3:.2f} M obj/s', flush=True)


if __name__ == '__main__':
    main()
