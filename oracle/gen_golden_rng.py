#!/usr/bin/env python
"""tests/golden/rng_stream.npz: where the UNMODIFIED reference leaves the global np.random stream.

TEST INFRASTRUCTURE.  Run in the build container only (`python oracle/gen_golden_rng.py`).
`getRANSACInliers` draws `np.random.randint(N, size=10)` inside its loop (PoseEst/pose_utils.py:73) and stops drawing
at the early-stop break (:80-81).  For each case the generator seeds the global stream, calls the real
`estimateSimilarityTransform(source, target, ratio_adapt=...)` WITHOUT patching the generator, and stores the outputs,
the number of iterations the loop ran (counted by a wrapper that delegates to the real `randint`) and the next four
values of the stream.  Cases: no early stop (100 iterations), a stop at the first iteration (exact data with identity
rotation: under SURVEY.md F3 only those score a near-zero residual), and stops in the middle (StopT moved into the
spread of the hypotheses' residuals through `ratio_adapt`).
"""
from __future__ import annotations

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


def run_reference(pu, seed, src, dst, ratio_adapt):
    """-> (scales, rotation, translation, transform, ok, iterations, next4)"""
    calls = {'n': 0}
    real = np.random.randint

    def counting(*a, **k):
        calls['n'] += 1
        return real(*a, **k)
    np.random.seed(seed)
    np.random.randint = counting
    try:
        with redirect_stdout(io.StringIO()):
            s, r, t, tf = pu.estimateSimilarityTransform(src, dst, ratio_adapt=ratio_adapt)
    finally:
        np.random.randint = real
    nxt = np.random.randint(2 ** 31 - 1, size=4)
    ok = s is not None
    if not ok:
        s, r, t, tf = np.zeros(3), np.zeros((3, 3)), np.zeros(3), np.zeros((4, 4))
    return np.asarray(s), np.asarray(r), np.asarray(t), np.asarray(tf), ok, calls['n'], nxt


def hypothesis_residuals(pu, seed, src, dst, n_iter=100):
    """Residual of every hypothesis the loop WOULD evaluate with this seed if it never stopped (same draws)."""
    np.random.seed(seed)
    sh, th = _hom(src), _hom(dst)
    res = []
    for _ in range(n_iter):
        idx = np.random.randint(src.shape[0], size=10)
        _, _, _, tf = pu.estimateSimilarityUmeyama(sh[:, idx], th[:, idx])
        r, _, _ = pu.evaluateModel(tf, sh, th, 1.0)
        res.append(r)
    return np.asarray(res)


def main():
    pu, _ = ref_import.load_reference()
    rng = np.random.default_rng(20260)
    out = {}
    names = []

    def add(name, seed, src, dst, ratio_adapt):
        s, r, t, tf, ok, iters, nxt = run_reference(pu, seed, src, dst, ratio_adapt)
        names.append(name)
        out[name + '_seed'] = np.int64(seed)
        out[name + '_src'] = src
        out[name + '_dst'] = dst
        out[name + '_ratio_adapt'] = np.float64(ratio_adapt)
        out[name + '_scales'], out[name + '_rotation'], out[name + '_translation'] = s, r, t
        out[name + '_transform'] = tf
        out[name + '_ok'] = np.bool_(ok)
        out[name + '_iterations'] = np.int64(iters)
        out[name + '_next'] = nxt
        print(f'{name}: N={src.shape[0]} ratio_adapt={ratio_adapt:.6g} iterations={iters} ok={ok} next={nxt[:2]}')

    # 1) ordinary object: noise + gross outliers, never below StopT
    n = 600
    src = rng.uniform(-0.5, 0.5, size=(n, 3))
    rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    rot *= np.sign(np.linalg.det(rot))
    dst = 1.3 * src @ rot.T + np.array([0.2, -0.1, -3.0]) + rng.normal(scale=0.01, size=(n, 3))
    dst[rng.random(n) < 0.1] += rng.uniform(8, 20, size=3)
    add('full', 101, src, dst, 1.0)

    # 2) exact data with identity rotation: the first hypothesis already scores ~1e-15 < StopT
    src = rng.uniform(-0.5, 0.5, size=(300, 3))
    dst = 1.7 * src + np.array([0.3, 0.2, -2.5])
    add('stop_first', 102, src, dst, 1.0)

    # 3) stops in the middle: noisy identity-rotation data, StopT placed inside the residual spread by ratio_adapt
    for j, (seed, q) in enumerate(((103, 0.08), (104, 0.03), (105, 0.15))):
        src = rng.uniform(-0.5, 0.5, size=(400, 3))
        dst = 1.2 * src + np.array([-0.2, 0.1, -3.5]) + rng.normal(scale=0.02, size=src.shape)
        while True:                                                # a seed whose first hypothesis is not the stopper
            res = hypothesis_residuals(pu, seed, src, dst)
            if int(np.argmax(res < np.quantile(res, q))) >= 1:
                break
            seed += 1000
        stop_t = np.quantile(res, q)
        t_bar = np.mean(np.linalg.norm(dst, axis=1))
        s_bar = np.mean(np.linalg.norm(src, axis=1))
        pass_1 = max(t_bar / s_bar, s_bar / t_bar)                 # PassT at ratio_adapt = 1 (pose_utils.py:91-95)
        first = int(np.argmax(res < stop_t))
        # keep StopT clear of every residual by a relative margin so that rounding cannot move the stop
        lo = res[res < stop_t].max()
        hi = res[res >= stop_t].min()
        stop_t = 0.5 * (lo + hi)
        assert (hi - lo) / hi > 1e-6 and first >= 1
        add(f'stop_mid_{j}', seed, src, dst, 100.0 * stop_t / pass_1)
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, 'rng_stream.npz'), **out)
    print('wrote', os.path.join(GOLD, 'rng_stream.npz'))


if __name__ == '__main__':
    main()
