"""Import the REAL reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box, so nothing that
runs there may call this; it is used by `oracle/gen_golden.py` (which writes the committed
fixtures under tests/golden/); bench.py's CPU arm imports the same modules from the copy that
`oracle/Makefile` stages under oracle/_ref/ (git-ignored; it travels to the GPU box).

`PoseEst/pose_estimation.py` imports open3d / matplotlib / detectron2 / easydict at module
scope (pose_estimation.py:5-12); none is installed here and none is touched by
`backproject`, `transform_pc`, `cam2world` or `sort_bbox`, so empty stand-in modules are
registered before the import.  `run_pose` itself cannot run (Open3D calls) -- see DESIGN.md.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')      # oracle/Makefile copies the files here


def _default_root() -> str:
    """/root/reference in the build container; on the GPU box (no reference tree) the copy oracle/Makefile staged."""
    if os.path.isfile(os.path.join('/root/reference', 'PoseEst', 'pose_utils.py')):
        return '/root/reference'
    return _STAGED


REFERENCE_ROOT = os.environ.get('POSEFIT_REFERENCE_ROOT') or _default_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'PoseEst', 'pose_utils.py'))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns (pose_utils, pose_estimation) -- the unmodified reference modules."""
    if not reference_available():
        raise FileNotFoundError(f'reference tree not found at {REFERENCE_ROOT}')
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ('open3d', 'matplotlib', 'matplotlib.pyplot', 'detectron2', 'detectron2.utils',
                 'detectron2.utils.visualizer', 'detectron2.structures', 'easydict'):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    sys.modules['detectron2.utils.visualizer'].__dict__.setdefault('GenericMask', object)
    sys.modules['detectron2.structures'].__dict__.setdefault('BoxMode', object)
    sys.modules['easydict'].__dict__.setdefault('EasyDict', dict)
    if 'matplotlib' in sys.modules and not hasattr(sys.modules['matplotlib'], 'pyplot'):
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    from PoseEst import pose_utils  # noqa: E402
    try:
        from PoseEst import pose_estimation  # noqa: E402
    except Exception:  # baseconfig may want attribute-style dicts
        class _AttrDict(dict):
            __getattr__ = dict.get
            __setattr__ = dict.__setitem__
        sys.modules['easydict'].EasyDict = _AttrDict
        for k in ('baseconfig', 'PoseEst.pose_estimation'):
            sys.modules.pop(k, None)
        from PoseEst import pose_estimation  # noqa: E402
    return pose_utils, pose_estimation


def inlier_indices_from_subset(full_hom: np.ndarray, subset_hom: np.ndarray) -> np.ndarray:
    """getRANSACInliers returns SourceHom[:, BestInlierIdx] (pose_utils.py:83), not the indices.
    BestInlierIdx is ascending, so a two-pointer scan over exact column equality recovers it."""
    idx, j = [], 0
    n_sub = subset_hom.shape[1]
    for i in range(full_hom.shape[1]):
        if j < n_sub and np.array_equal(full_hom[:, i], subset_hom[:, j]):
            idx.append(i)
            j += 1
    assert j == n_sub, 'subset is not an ordered sub-sequence'
    return np.asarray(idx, dtype=np.int64)


@contextlib.contextmanager
def replay_randint(sample_idx: np.ndarray, pose_utils=None):
    """Patch np.random.randint so getRANSACInliers (pose_utils.py:73) draws `sample_idx[h]`
    on its h-th call, then restore it.  estimateSimilarityTransform hard-codes nIter = 100
    (pose_utils.py:97); when `pose_utils` is given, its getRANSACInliers is wrapped so that
    MaxIterations = len(sample_idx) -- the loop body that runs is still the reference's."""
    calls = {'n': 0}
    real = np.random.randint

    def fake(high, size=None, **kw):
        h = calls['n']
        calls['n'] += 1
        row = np.asarray(sample_idx[h])
        assert size is None or row.shape[0] == (size if np.isscalar(size) else size[0])
        assert row.max(initial=0) < high
        return row.copy()

    real_ransac = getattr(pose_utils, 'getRANSACInliers', None)
    if real_ransac is not None:
        def capped(SourceHom, TargetHom, MaxIterations=100, PassThreshold=200, StopThreshold=1):
            ret = real_ransac(SourceHom, TargetHom, MaxIterations=len(sample_idx),
                              PassThreshold=PassThreshold, StopThreshold=StopThreshold)
            calls['ransac_out'] = ret          # (SourceInliersHom, TargetInliersHom, BestInlierRatio)
            calls['pass_t'] = PassThreshold
            return ret
        pose_utils.getRANSACInliers = capped
    np.random.randint = fake
    try:
        yield calls
    finally:
        np.random.randint = real
        if real_ransac is not None:
            pose_utils.getRANSACInliers = real_ransac
