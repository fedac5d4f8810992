"""csrc/posefit_math.h (the per-object arithmetic of the CUDA kernels) compiled for the host
and pinned to the oracle and the golden vectors -- no GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import grad_oracle
from oracle import posefit_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, '3d_mot_differentiable_pose_estimation_b200', 'csrc')


@pytest.fixture(scope='module')
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp('host') / 'math_check.so'
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', f'-I{CSRC}',
                           os.path.join(ROOT, 'tests', 'host', 'math_check.cpp'), '-o', str(out)])
    return ctypes.CDLL(str(out))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def c_fit(lib, src, dst, precise=True):
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    out = np.zeros(27)
    lib.pf_check_fit(_dp(src), _dp(dst), ctypes.c_int(src.shape[0]), ctypes.c_int(int(precise)), _dp(out))
    return dict(s=out[0], R=out[1:10].reshape(3, 3), t=out[10:13], var=out[13], H=out[14:20], Linv=out[20:26],
                status=int(out[26]))


def rot_angle_deg(ra, rb):
    c = (np.trace(ra.T @ rb) - 1) / 2
    return np.degrees(np.arccos(np.clip(c, -1, 1)))


def rot_err_deg(ra, rb):
    # robust for tiny angles: |Ra - Rb|_F / sqrt(2) ~ angle
    return np.degrees(np.linalg.norm(ra - rb) / np.sqrt(2))


@pytest.mark.parametrize('precise', [True, False])
def test_fit_matches_golden_umeyama(lib, golden_dir, precise):
    g = np.load(os.path.join(golden_dir, 'umeyama.npz'))
    for k, name in enumerate(str(n) for n in g['names']):
        if name in ('roundoff_variance', 'single_point'):
            continue        # chaotic / rank-0 cases: the reference's own answer is rounding noise
        f = c_fit(lib, g[f'src_{k}'], g[f'dst_{k}'], precise)
        assert f['status'] == 0
        ref_R = g[f'rotation_{k}'].T
        assert rot_err_deg(f['R'], ref_R) < 1e-7, name
        np.testing.assert_allclose(f['s'], g[f'scales_{k}'][0], rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(f['t'], g[f'translation_{k}'], rtol=1e-9, atol=1e-9, err_msg=name)


def test_fit_degenerate_cases(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, 'umeyama.npz'))
    names = [str(n) for n in g['names']]
    k = names.index('zero_variance')
    f = c_fit(lib, g[f'src_{k}'], g[f'dst_{k}'])
    assert f['s'] == 1.0 and np.array_equal(f['R'], np.identity(3))          # pose_utils.py:47-50
    np.testing.assert_allclose(f['t'], g[f'translation_{k}'], rtol=1e-14)
    k = names.index('single_point')
    f = c_fit(lib, g[f'src_{k}'], g[f'dst_{k}'])
    assert f['status'] == 0 and f['s'] == 1.0
    f = c_fit(lib, np.zeros((0, 3)), np.zeros((0, 3)))
    assert f['status'] == 1
    bad = np.zeros((5, 3))
    bad[2, 1] = np.nan
    assert c_fit(lib, bad, np.ones((5, 3)))['status'] == 3                   # pose_utils.py:32-36


@pytest.mark.parametrize('precise', [True, False])
def test_fit_random_sweep_against_oracle(lib, precise):
    rng = np.random.default_rng(7)
    worst = 0.0
    for trial in range(400):
        n = int(rng.choice([3, 4, 10, 10, 10, 50, 500]))
        src = rng.uniform(-0.5, 0.5, size=(n, 3))
        kind = trial % 5
        if kind == 1:
            src[:, 2] *= 10.0 ** -rng.integers(1, 6)                # thin slab
        q = rng.normal(size=(3, 3))
        rot = np.linalg.qr(q)[0]
        if np.linalg.det(rot) < 0:
            rot[:, 0] = -rot[:, 0]
        dst = rng.uniform(0.3, 3) * src @ rot.T + rng.uniform(-4, 4, size=3)
        dst += rng.normal(scale=10.0 ** -rng.integers(1, 5), size=dst.shape)
        if kind == 2:
            dst[:, 0] = -dst[:, 0]                                   # forces the reflection branch often
        if kind == 3:
            dst = rng.uniform(-4, 4, size=(n, 3))                    # unrelated clouds
        scales, rot_t, trans, _ = po.umeyama_fit(src, dst)
        # conditioning of the rotation: smallest h_i + h_j of the sign-fixed singular values
        cov = (dst - dst.mean(0)).T @ (src - src.mean(0)) / n
        u, d, vh = np.linalg.svd(cov)
        if np.linalg.det(u) * np.linalg.det(vh) < 0:
            d[2] = -d[2]
        gap = (d[1] + d[2]) / d[0]
        f = c_fit(lib, src, dst, precise)
        assert f['status'] == 0
        err = rot_err_deg(f['R'], rot_t.T)
        if gap > 1e-6:
            assert err < 1e-6 / gap * 1e-3 + 1e-9, (trial, n, kind, gap, err)
            np.testing.assert_allclose(f['s'], scales[0], rtol=1e-8 / min(gap, 1.0) * 1e-3 + 1e-10)
            worst = max(worst, err)
        assert abs(np.linalg.det(f['R']) - 1) < 1e-12
        np.testing.assert_allclose(f['R'] @ f['R'].T, np.identity(3), atol=1e-12)


def test_closed_form_residuals_match_score_model(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, 'ransac.npz'))
    for k in range(int(g['count'])):
        src, dst, idx = g[f'src_{k}'], g[f'dst_{k}'], g[f'idx_{k}'].astype(np.int32)
        src = np.ascontiguousarray(src)
        dst = np.ascontiguousarray(dst)
        idx = np.ascontiguousarray(idx)
        res2 = np.zeros(idx.shape[0])
        lib.pf_check_hypotheses(_dp(src), _dp(dst), ctypes.c_int(src.shape[0]),
                                idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.c_int(idx.shape[0]),
                                ctypes.c_int(idx.shape[1]), ctypes.c_int(1), _dp(res2))
        want = np.array([po.score_model(po.umeyama_fit(src[i], dst[i])[3], src, dst, 1.0)[0] for i in idx])
        # closed form vs direct sum: relative 1e-12, plus cancellation noise ~1e-15 * sum|y~|^2 on exact fits
        scale = float(((dst - dst.mean(0)) ** 2).sum())
        np.testing.assert_allclose(res2, want ** 2, rtol=1e-12, atol=1e-14 * scale, err_msg=str(g[f'name_{k}']))
        if want.min() > 1e-6:
            assert np.argmin(res2) == np.argmin(want)


def test_adjoint_matches_autograd(lib):
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.choice([4, 10, 200]))
        src = rng.uniform(-0.5, 0.5, size=(n, 3))
        rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        dst = rng.uniform(0.5, 2.5) * src @ rot.T + rng.uniform(-3, 3, size=3) + rng.normal(scale=0.05, size=(n, 3))
        gs, gR, gt = rng.normal(), rng.normal(size=(3, 3)), rng.normal(size=3)
        gx, gy, (s, R, t) = grad_oracle.fit_gradients(torch.from_numpy(src), torch.from_numpy(dst), None,
                                                      gs, torch.from_numpy(gR), torch.from_numpy(gt))
        # the torch transcription's forward equals the NumPy oracle
        scales, rot_t, trans, _ = po.umeyama_fit(src, dst)
        np.testing.assert_allclose(s.numpy(), scales[0], rtol=1e-12)
        np.testing.assert_allclose(R.numpy(), rot_t.T, atol=1e-12)
        np.testing.assert_allclose(t.numpy(), trans, atol=1e-12)
        out = np.zeros(16)
        srcc, dstc, gRc, gtc = (np.ascontiguousarray(a) for a in (src, dst, gR, gt))
        lib.pf_check_adjoint(_dp(srcc), _dp(dstc), ctypes.c_int(n), ctypes.c_double(gs), _dp(gRc), _dp(gtc), _dp(out))
        GC, gvar, gmux, gmuy = out[:9].reshape(3, 3), out[9], out[10:13], out[13:16]
        xt, yt = src - src.mean(0), dst - dst.mean(0)
        my_gx = (yt @ GC + 2 * gvar * xt + gmux) / n
        my_gy = (xt @ GC.T + gmuy) / n
        np.testing.assert_allclose(my_gx, gx.numpy(), rtol=1e-8, atol=1e-10 * np.abs(gx.numpy()).max())
        np.testing.assert_allclose(my_gy, gy.numpy(), rtol=1e-8, atol=1e-10 * np.abs(gy.numpy()).max())


def _fd_case(g, k):
    """One case of tests/golden/grad_fd.npz -> correspondences in np.where order + the FD gradients at them."""
    noc, depth, mask, xy0 = g[f'noc_{k}'], g[f'depth_{k}'], g[f'mask_{k}'], g[f'xy0_{k}']
    h, w = depth.shape
    x0, y0 = int(xy0[0]), int(xy0[1])
    frame_d = np.zeros((po.FRAME_H, po.FRAME_W), dtype=np.float32)
    frame_m = np.zeros((po.FRAME_H, po.FRAME_W), dtype=bool)
    frame_d[y0:y0 + h, x0:x0 + w] = depth
    frame_m[y0:y0 + h, x0:x0 + w] = mask != 0
    k_mat = po.motfront_intrinsics()
    noc_pts, depth_pts, (rows, cols) = po.crop_correspondences(np.transpose(noc, (1, 2, 0)), frame_d, frame_m,
                                                               (x0, y0, x0 + w, y0 + h), k_mat)
    want_x = g[f'grad_noc_{k}'][:, rows - y0, cols - x0].T               # [N,3]
    want_z = g[f'grad_depth_{k}'][rows - y0, cols - x0]                  # [N]
    rx = (cols - k_mat[0, 2]) / k_mat[0, 0]
    ry = (rows - k_mat[1, 2]) / k_mat[1, 1]
    return noc_pts, depth_pts, want_x, want_z, rx, ry


def test_gradients_match_reference_finite_differences(lib, golden_dir):
    """PINS the gradient path: tests/golden/grad_fd.npz holds central finite differences of the REAL
    reference functions (backproject + estimateSimilarityUmeyama, oracle/gen_golden_grad.py).  Both the
    autograd restatement and the host-compiled adjoint of csrc/posefit_math.h must reproduce them
    (FD accuracy ~1e-8 relative: tolerance 2e-6 of the largest component)."""
    g = np.load(os.path.join(golden_dir, 'grad_fd.npz'))
    for k in range(int(g['n_cases'])):
        noc_pts, depth_pts, want_x, want_z, rx, ry = _fd_case(g, k)
        n = noc_pts.shape[0]
        assert n == int(g[f'n_valid_{k}'])
        gs, gR, gt = float(g[f'g_s_{k}']), g[f'g_R_{k}'], g[f'g_t_{k}']
        gx, gy, (s, R, t) = grad_oracle.fit_gradients(torch.from_numpy(noc_pts), torch.from_numpy(depth_pts), None,
                                                      gs, torch.from_numpy(gR), torch.from_numpy(gt))
        np.testing.assert_allclose(float(s), float(g[f's_{k}']), rtol=1e-12)
        np.testing.assert_allclose(R.numpy(), g[f'R_{k}'], atol=1e-12)
        out = np.zeros(16)
        srcc, dstc, gRc, gtc = (np.ascontiguousarray(a) for a in (noc_pts, depth_pts, gR, gt))
        lib.pf_check_adjoint(_dp(srcc), _dp(dstc), ctypes.c_int(n), ctypes.c_double(gs), _dp(gRc), _dp(gtc), _dp(out))
        GC, gvar, gmux, gmuy = out[:9].reshape(3, 3), out[9], out[10:13], out[13:16]
        xt, yt = noc_pts - noc_pts.mean(0), depth_pts - depth_pts.mean(0)
        for name, got_x, got_y in (('autograd restatement', gx.numpy(), gy.numpy()),
                                   ('posefit_math.h adjoint', (yt @ GC + 2 * gvar * xt + gmux) / n, (xt @ GC.T + gmuy) / n)):
            got_z = got_y[:, 0] * rx - got_y[:, 1] * ry - got_y[:, 2]    # y = (rx z, -ry z, -z), pose_estimation.py:34-41
            assert np.abs(got_x - want_x).max() <= 2e-6 * np.abs(want_x).max(), (name, k)
            assert np.abs(got_z - want_z).max() <= 2e-6 * np.abs(want_z).max(), (name, k)


def test_ransac_gradients_match_reference_finite_differences(golden_dir):
    """Same pin for the RANSAC path: finite differences of the real estimateSimilarityTransform (replayed
    indices) vs the autograd restatement weighted by the oracle's inlier set (winner and inliers locally constant)."""
    g = np.load(os.path.join(golden_dir, 'grad_fd.npz'))
    k = int(g['ransac_case'])
    noc_pts, depth_pts, want_x, want_z, rx, ry = _fd_case(g, k)
    o = po.similarity_transform(noc_pts, depth_pts, g[f'sample_idx_{k}'])
    assert o['ok'] and 0 < len(o['inlier_idx']) < noc_pts.shape[0]
    wts = np.zeros(noc_pts.shape[0])
    wts[o['inlier_idx']] = 1.0
    gx, gy, (s, R, t) = grad_oracle.fit_gradients(torch.from_numpy(noc_pts), torch.from_numpy(depth_pts),
                                                  torch.from_numpy(wts), float(g[f'g_s_{k}']),
                                                  torch.from_numpy(g[f'g_R_{k}']), torch.from_numpy(g[f'g_t_{k}']))
    np.testing.assert_allclose(float(s), float(g[f's_{k}']), rtol=1e-12)
    got_x, got_y = gx.numpy(), gy.numpy()
    got_z = got_y[:, 0] * rx - got_y[:, 1] * ry - got_y[:, 2]
    assert np.abs(got_x - want_x).max() <= 2e-6 * np.abs(want_x).max()
    assert np.abs(got_z - want_z).max() <= 2e-6 * np.abs(want_z).max()
    assert np.all(want_x[wts == 0] == 0.0)                           # outliers of the winner: exactly no gradient


# ---------------------------------------------------------------------------------------------
# float screen of the RANSAC hypotheses (fit_ransac_kernel<.., SCREEN>): the residual interval of the float fit
# must contain the double fit's residual, and the candidate rule must keep the reference's winner
# ---------------------------------------------------------------------------------------------
SCREEN_REGIMES = [
    ('bench', dict(n=96, h=64, w=64, n_hyp=128, n_samp=10)),
    ('clean', dict(n=48, h=64, w=64, n_hyp=128, n_samp=10, outlier_frac=0.0)),
    ('no_noise', dict(n=24, h=64, w=64, n_hyp=64, n_samp=10, outlier_frac=0.0, noc_noise=0.0)),
    ('heavy', dict(n=48, h=64, w=64, n_hyp=128, n_samp=10, outlier_frac=0.4)),
    ('sparse', dict(n=128, h=24, w=28, n_hyp=100, n_samp=10, mask_fill=0.3)),
    ('tiny3', dict(n=256, h=12, w=12, n_hyp=32, n_samp=3, mask_fill=0.3, border=0)),
    ('tiny16', dict(n=128, h=16, w=20, n_hyp=64, n_samp=16, mask_fill=0.1, border=0)),
    ('large', dict(n=12, h=112, w=112, n_hyp=128, n_samp=10)),
]


@pytest.mark.parametrize('name,cfg', SCREEN_REGIMES, ids=[r[0] for r in SCREEN_REGIMES])
def test_screen_interval_contains_double_residual(lib, name, cfg):
    import importlib
    import zlib
    synth = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.synth')
    cfg = dict(cfg)
    n_obj, h, w = cfg.pop('n'), cfg.pop('h'), cfg.pop('w')
    d = synth.make_objects(n_obj, h, w, seed=zlib.crc32(name.encode()) % 10000 + 7, **cfg)
    noc, depth, mask = d['noc'].numpy(), d['depth'].numpy(), d['mask'].numpy()
    xy0, idx = d['bbox_xy0'].numpy(), d['sample_idx'].numpy()
    kinv = np.linalg.inv(po.motfront_intrinsics())
    kinv4 = np.array([kinv[0, 0], kinv[0, 2], kinv[1, 1], kinv[1, 2]])
    n_hyp, n_samp = idx.shape[1], idx.shape[2]
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    n_fin = n_cand = n_objs = 0
    worst = 0.0
    for i in range(n_obj):
        rows, cols = np.where((mask[i] != 0) & (depth[i] > 0))
        n = rows.size
        if n == 0:
            continue
        pts_noc = np.ascontiguousarray(noc[i][:, rows, cols].T, dtype=np.float32)
        z = np.ascontiguousarray(depth[i][rows, cols], dtype=np.float32)
        fr = np.ascontiguousarray(rows + xy0[i, 1], dtype=np.int32)
        fc = np.ascontiguousarray(cols + xy0[i, 0], dtype=np.int32)
        id_ = np.ascontiguousarray(idx[i], dtype=np.int32)
        r64, lo, hi = np.zeros(n_hyp), np.zeros(n_hyp), np.zeros(n_hyp)
        lib.pf_check_screen(fp(pts_noc), fp(z), ip(fr), ip(fc), ctypes.c_int(n), _dp(kinv4), ip(id_),
                            ctypes.c_int(n_hyp), ctypes.c_int(n_samp), ctypes.c_int(1), _dp(r64), _dp(lo), _dp(hi))
        fin = np.isfinite(lo) & np.isfinite(hi) & np.isfinite(r64)
        assert np.all(lo[fin] <= r64[fin]) and np.all(r64[fin] <= hi[fin]), (name, i)
        half = 0.5 * (hi[fin] - lo[fin])
        if fin.any():
            worst = max(worst, float(np.max(np.abs(r64[fin] - 0.5 * (hi[fin] + lo[fin])) / np.maximum(half, 1e-300))))
        # candidate rule of the kernel: keeps the first minimum (and every hypothesis below the stop threshold)
        u = np.min(np.where(np.isnan(hi), np.inf, hi))
        cand = ~(lo > u)
        if np.isfinite(r64).any():
            win = int(np.nanargmin(r64))
            assert cand[win], (name, i, win)
        n_fin += int(fin.sum())
        n_cand += int(cand.sum())
        n_objs += 1
    assert n_objs > 0
    print(f'screen[{name}]: {n_objs} objects, {n_fin} finite intervals, worst |error| / half-width = {worst:.3g}, '
          f'candidates per object = {n_cand / n_objs:.2f}')
    assert worst < 0.5          # the interval is at least twice as wide as any error seen
