"""Drop-in for the reference's `PoseEst/pose_utils.py`: same function names, arguments, return
types and error behaviour, computed by the CUDA library (points-mode entries of include/posefit.h).

These are per-object compatibility wrappers: they move one cloud to the GPU, call the batched
kernels with B = 1 and hand NumPy arrays back, exactly like the reference's functions.  The
batched, sync-free operator is `function.PoseFit` / `pose_fit_raw`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .function import points_fit_raw, ransac_iterations, _ptr, _stream

__all__ = ['evaluateModel', 'estimateSimilarityUmeyama', 'getRANSACInliers', 'estimateSimilarityTransform']

N_SAMPLES = 10          # pose_utils.py:73
N_ITERATIONS = 100      # pose_utils.py:97


def _device():
    if not torch.cuda.is_available():
        raise _lib.PoseFitError('pose_utils needs a CUDA device: the solver has no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def _planes(hom) -> torch.Tensor:
    """[4,N] (or [3,N]) homogeneous array -> [1,3,N] float64 CUDA tensor (rows 0..2)."""
    a = np.asarray(hom, dtype=np.float64)[:3, :]
    return torch.from_numpy(np.ascontiguousarray(a)).to(_device())[None]


def _as_reference(pose_row: np.ndarray):
    """pose record -> (Scales, Rotation, Translation, OutTransform) of pose_utils.py:52-61."""
    s = pose_row[0]
    rot_t = pose_row[1:10].reshape(3, 3).T.copy()        # the reference reports R^T (:44)
    trans = pose_row[10:13].copy()
    scales = np.array([s, s, s])
    out = np.identity(4)
    out[:3, :3] = np.diag(scales) @ rot_t                # :58
    out[:3, 3] = trans
    return scales, rot_t, trans, out


def estimateSimilarityUmeyama(SourceHom, TargetHom):
    """pose_utils.py:16-61.  Raises RuntimeError('There are NANs in the input.') like :32-36."""
    src, dst = _planes(SourceHom), _planes(TargetHom)
    raw = points_fit_raw(src, dst)
    status = int(raw.status[0])
    if status == 3:
        print('nPoints:', src.shape[2])
        raise RuntimeError('There are NANs in the input.')
    return _as_reference(raw.pose[0].cpu().numpy())


def evaluateModel(OutTransform, SourceHom, TargetHom, PassThreshold):
    """pose_utils.py:5-14 -> (Residual, InlierRatio, InlierIdx)."""
    lib = _lib.lib()
    src, dst = _planes(SourceHom), _planes(TargetHom)
    dev = src.device
    n = int(src.shape[2])
    tf = torch.from_numpy(np.ascontiguousarray(np.asarray(OutTransform, dtype=np.float64).reshape(1, 16))).to(dev)
    mask = torch.ones(1, n, dtype=torch.uint8, device=dev)
    pt = torch.tensor([float(PassThreshold)], dtype=torch.float64, device=dev)
    stats = torch.empty(1, 4, dtype=torch.float64, device=dev)
    inl = torch.empty(1, n, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_points_evaluate(_ptr(tf), _ptr(src), _ptr(dst), _ptr(mask), _ptr(pt), 0, 1, n,
                                           _ptr(stats), _ptr(inl), _stream(dev))
    _lib.check(code, 'posefit_points_evaluate')
    st = stats[0].cpu().numpy()
    idx = np.nonzero(inl[0].cpu().numpy())[0]
    n_counted = st[1] - st[2]                             # count_nonzero skips index 0 (:11)
    return float(st[0]), float(n_counted / n), idx


def _ransac(SourceHom, TargetHom, n_iter, pass_t, stop_t, ratio_adapt=1.0):
    src, dst = _planes(SourceHom), _planes(TargetHom)
    n = int(src.shape[2])
    state = np.random.get_state()
    idx = np.random.randint(n, size=(n_iter, N_SAMPLES))  # :73 -- drawn up front from the same global RNG
    raw = points_fit_raw(src, dst, sample_idx=torch.from_numpy(idx.astype(np.int32))[None], ratio_adapt=ratio_adapt,
                         pass_threshold=pass_t, stop_threshold=stop_t)
    rewind_to_reference_stream(state, n, int(ransac_iterations(raw)[0]), n_iter)
    return raw


def rewind_to_reference_stream(state, n_points: int, iterations_run: int, n_iter: int, n_samples: int = N_SAMPLES):
    """The reference draws `np.random.randint(N, size=10)` INSIDE its loop and stops drawing when the loop breaks
    (pose_utils.py:73, :80-81); the drop-ins draw every iteration's indices up front.  After an early stop put the global
    stream where the reference leaves it: back to `state` (taken before the draw), then exactly the draws of the
    iterations that ran.  (Legacy `randint` produces its values one after another from the bit stream, so one call of
    `iterations_run` x 10 consumes what `iterations_run` calls of 10 consume; tests/golden/rng_stream.npz pins it.)"""
    if 0 <= iterations_run < n_iter:
        np.random.set_state(state)
        if iterations_run > 0:
            np.random.randint(n_points, size=(iterations_run, n_samples))


def getRANSACInliers(SourceHom, TargetHom, MaxIterations=100, PassThreshold=200, StopThreshold=1):
    """pose_utils.py:63-83 -> (SourceHom[:, inliers], TargetHom[:, inliers], BestInlierRatio)."""
    raw = _ransac(SourceHom, TargetHom, int(MaxIterations), float(PassThreshold), float(StopThreshold))
    keep = np.nonzero(raw.inlier_mask[0].cpu().numpy())[0]
    ratio = float(raw.pose[0, 14])
    SourceHom, TargetHom = np.asarray(SourceHom), np.asarray(TargetHom)
    return SourceHom[:, keep], TargetHom[:, keep], ratio


def estimateSimilarityTransform(source: np.array, target: np.array, verbose=False, ratio_adapt=1):
    """pose_utils.py:86-117 -> (Scales, Rotation, Translation, OutTransform) or 4 x None."""
    source, target = np.asarray(source, dtype=np.float64), np.asarray(target, dtype=np.float64)
    raw = _ransac(source.T, target.T, N_ITERATIONS, 0.0, 0.0, float(ratio_adapt))
    pose = raw.pose[0].cpu().numpy()
    status = int(raw.status[0])
    if verbose:
        print('Pass threshold: ', pose[15])
        print('Stop threshold: ', pose[15] / 100)
        print('Number of iterations: ', N_ITERATIONS)
    if status == 2:
        print('[ WARN ] - Something is wrong. Small BestInlierRatio: ', pose[14])
        return None, None, None, None
    if status == 3:
        raise RuntimeError('There are NANs in the input.')
    scales, rot_t, trans, out = _as_reference(pose)
    if verbose:
        print('BestInlierRatio:', pose[14])
        print('Rotation:\n', rot_t)
        print('Translation:\n', trans)
        print('Scales:', scales)
    return scales, rot_t, trans, out
