"""Guard-zone test of the C ABI (compute-sanitizer is not available on this pool): every output and the
workspace of posefit_forward / posefit_forward_ransac / posefit_backward is carved out of one arena filled
with a sentinel, with 256-byte guard zones between the pieces.  After the calls the guard zones must be
untouched (no out-of-bounds write for any crop shape / batch size), the outputs fully written, and the
workspace sizes the library reports sufficient."""
import numpy as np
import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.gpu

GUARD = 256
SENTINEL = 0xA5


@pytest.fixture(scope='module')
def pf():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return load_pkg()


class Arena:
    """One device buffer; pieces are 256-byte aligned and separated by sentinel guard zones."""

    def __init__(self, nbytes):
        self.buf = torch.full((nbytes,), SENTINEL, dtype=torch.uint8, device='cuda')
        self.off = GUARD
        self.pieces = []

    def take(self, nbytes):
        start = self.off
        self.off = (start + nbytes + GUARD + 255) // 256 * 256
        assert self.off <= self.buf.numel()
        self.pieces.append((start, nbytes))
        return self.buf.data_ptr() + start, self.buf[start:start + nbytes]

    def guards_intact(self):
        keep = torch.ones(self.buf.numel(), dtype=torch.bool, device='cuda')
        for start, n in self.pieces:
            keep[start:start + n] = False
        return bool((self.buf[keep] == SENTINEL).all())


@pytest.mark.parametrize('h,w,b,n_hyp', [(64, 64, 37, 128), (3, 5, 4, 0), (17, 16, 9, 12), (20, 20, 300, 16),
                                         (112, 112, 5, 32), (33, 64, 11, 130), (150, 172, 2, 8), (1, 1, 3, 0),
                                         (64, 64, 4096, 0), (16, 16, 4737, 8), (16, 16, 149, 8)])
def test_no_write_outside_outputs(pf, h, w, b, n_hyp):
    lib = pf._lib.lib()
    d = pf.synth.make_objects(b, h, w, seed=5, device='cuda', n_hyp=max(n_hyp, 1), align_x0=1 if w % 4 else 4)
    kinv = pf.default_kinv('cuda').contiguous()
    p = h * w
    stream = torch.cuda.current_stream().cuda_stream
    ws_plain = int(lib.posefit_workspace_bytes(b, h, w, 0, 0))
    ws_ransac = int(lib.posefit_workspace_bytes(b, h, w, max(n_hyp, 1), 10))
    ws_bwd = int(lib.posefit_backward_workspace_bytes(b))
    sizes = dict(pose=b * 16 * 8, ctx=b * 32 * 8, status=b * 4, n_valid=b * 4, inl=b * p, winner=b * 4,
                 g_noc=b * 3 * p * 4, g_depth=b * p * 4, ws=max(ws_plain, ws_ransac, ws_bwd, 16))
    arena = Arena(sum(v + 2 * GUARD + 256 for v in sizes.values()) + 4096)
    ptr, view = {}, {}
    for k, v in sizes.items():
        ptr[k], view[k] = arena.take(v)
    noc, depth, mask, xy0 = d['noc'], d['depth'], d['mask'], d['bbox_xy0']
    code = lib.posefit_forward(noc.data_ptr(), depth.data_ptr(), mask.data_ptr(), xy0.data_ptr(), kinv.data_ptr(), 0,
                               b, h, w, ptr['pose'], ptr['ctx'], ptr['status'], ptr['n_valid'], ptr['ws'], ws_plain,
                               stream)
    assert code == 0, lib.posefit_error_string(code)
    torch.cuda.synchronize()
    assert arena.guards_intact(), 'posefit_forward wrote outside its outputs / workspace'
    status = view['status'].view(torch.int32)
    assert int(((status < 0) | (status > 3)).sum()) == 0                 # fully written (sentinel would be 0xA5A5A5A5)
    gs = torch.ones(b, dtype=torch.float32, device='cuda')
    gR = torch.ones(b, 9, dtype=torch.float32, device='cuda')
    gt = torch.ones(b, 3, dtype=torch.float32, device='cuda')
    code = lib.posefit_backward(noc.data_ptr(), depth.data_ptr(), mask.data_ptr(), None, xy0.data_ptr(), kinv.data_ptr(),
                                0, b, h, w, ptr['ctx'], ptr['status'], gs.data_ptr(), gR.data_ptr(), gt.data_ptr(),
                                ptr['g_noc'], ptr['g_depth'], ptr['ws'], ws_bwd, stream)
    assert code == 0, lib.posefit_error_string(code)
    torch.cuda.synchronize()
    assert arena.guards_intact(), 'posefit_backward wrote outside its outputs / workspace'
    assert torch.isfinite(view['g_noc'].view(torch.float32)).all() and torch.isfinite(view['g_depth'].view(torch.float32)).all()
    if n_hyp:
        idx = d['sample_idx'].contiguous()
        code = lib.posefit_forward_ransac(noc.data_ptr(), depth.data_ptr(), mask.data_ptr(), xy0.data_ptr(),
                                          kinv.data_ptr(), 0, idx.data_ptr(), b, h, w, n_hyp, 10, 1.0, 1, ptr['pose'],
                                          ptr['ctx'], ptr['status'], ptr['n_valid'], ptr['inl'], ptr['winner'], ptr['ws'],
                                          ws_ransac, stream)
        assert code == 0, lib.posefit_error_string(code)
        torch.cuda.synchronize()
        assert arena.guards_intact(), 'posefit_forward_ransac wrote outside its outputs / workspace'
        assert int((view['inl'] > 1).sum()) == 0                         # every mask byte written (0 / 1)
        code = lib.posefit_backward(noc.data_ptr(), depth.data_ptr(), mask.data_ptr(), ptr['inl'], xy0.data_ptr(),
                                    kinv.data_ptr(), 0, b, h, w, ptr['ctx'], ptr['status'], gs.data_ptr(), gR.data_ptr(),
                                    gt.data_ptr(), ptr['g_noc'], ptr['g_depth'], ptr['ws'], ws_bwd, stream)
        assert code == 0, lib.posefit_error_string(code)
        torch.cuda.synchronize()
        assert arena.guards_intact()
    # a workspace one byte short is refused, not overrun
    if ws_plain > 8:
        code = lib.posefit_forward(noc.data_ptr(), depth.data_ptr(), mask.data_ptr(), xy0.data_ptr(), kinv.data_ptr(), 0,
                                   b, h, w, ptr['pose'], ptr['ctx'], ptr['status'], ptr['n_valid'], ptr['ws'],
                                   ws_plain - 1, stream)
        assert code != 0
    np.testing.assert_equal(arena.guards_intact(), True)


def test_bounds_checked_build_runs_clean():
    """compute-sanitizer is closed on this pool; the -DPF_BOUNDS build of the library (csrc/Makefile: bounds) checks every
    computed index of the three hot kernels -- shared-memory rings and tables, bitmap, select list, queues, per-object
    offsets -- and traps on a violation.  A representative workload (plain fit, both RANSAC kernels with several
    objects per CTA, backward, ragged shapes, sparse masks) must run through it without an error, and give the same
    results as the product build."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, '3d_mot_differentiable_pose_estimation_b200', 'libposefit_b200_bounds.so')
    if not os.path.exists(lib):
        pytest.skip('bounds build not present (make -C csrc bounds)')
    code = r"""
import importlib, sys, torch
pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
out = []
for (b, h, w, kw) in [(1500, 64, 64, {}), (40, 112, 112, {}), (64, 19, 27, dict(align_x0=1)), (300, 32, 32, dict(mask_fill=0.1, border=0)),
                      (64, 48, 64, dict(outlier_frac=0.4)), (16, 64, 64, dict(outlier_frac=0.0, noc_noise=0.0))]:
    d = pf.synth.make_objects(b, h, w, seed=7, n_hyp=64, device='cuda', **kw)
    g = (torch.ones(b, device='cuda'), torch.ones(b, 9, device='cuda'), torch.ones(b, 3, device='cuda'))
    raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'])
    rr = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], sample_idx=d['sample_idx'])
    gn, _ = pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], rr.inlier_mask, d['bbox_xy0'], None, rr.ctx, rr.status, *g)
    torch.cuda.synchronize()
    out.append((float(raw.pose[:, :13].sum()), float(rr.pose[:, :13].sum()), int(rr.inlier_mask.sum()), float(gn.double().abs().sum())))
print(repr(out))
"""
    results = []
    for which in (lib, ''):
        env = dict(os.environ)
        env.pop('POSEFIT_LIB', None)
        if which:
            env['POSEFIT_LIB'] = which
        res = subprocess.run([sys.executable, '-c', code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        results.append(res.stdout.strip().splitlines()[-1])
    assert results[0] == results[1], results
