#!/usr/bin/env python
"""Golden vectors for sort_bbox (PoseEst/pose_estimation.py:72-93), produced by the UNMODIFIED reference function.

TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference).  Writes tests/golden/sort_bbox.npz:
  in_k / out_k   8x3 corner sets and what the real sort_bbox returns for them
  tag_k          what the case is
Cases: axis-aligned boxes with their corners in the order the product's epilogue hands them to sort_bbox
(pose_estimation.py:373-380 takes Open3D's AxisAlignedBoundingBox.get_box_points(); that order is restated in
pose_estimation._aabb_corners and stays UNPINNED -- Open3D is not installed), for every combination of zero / positive
extent per axis (argsort ties!), with negative and mixed-sign coordinates; the same corners shuffled; arbitrary 8x3
arrays; arrays with repeated rows and repeated single coordinates."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def aabb_corners(lo, hi):
    """The corner order pose_estimation._aabb_corners documents (Open3D get_box_points, unpinned)."""
    e = hi - lo
    return np.array([lo, lo + [e[0], 0, 0], lo + [0, e[1], 0], lo + [0, 0, e[2]],
                     hi, hi - [e[0], 0, 0], hi - [0, e[1], 0], hi - [0, 0, e[2]]])


def main():
    _, pe = ref_import.load_reference()
    rng = np.random.default_rng(20260318)
    out = {}
    k = 0

    def add(tag, pts):
        nonlocal k
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out[f'in_{k}'] = pts
        out[f'out_{k}'] = np.asarray(pe.sort_bbox(pts.copy()), dtype=np.float64)
        out[f'tag_{k}'] = np.array(tag)
        k += 1

    for pattern in range(8):                                   # bit a set: positive extent along axis a
        for rep in range(6):
            lo = rng.uniform(-4, 4, size=3)
            ext = np.array([(pattern >> a) & 1 for a in range(3)]) * rng.uniform(0.1, 3.0, size=3)
            c = aabb_corners(lo, lo + ext)
            add(f'aabb extent mask {pattern}', c)
            add(f'aabb extent mask {pattern} shuffled', c[rng.permutation(8)])
    for rep in range(24):
        add('random 8x3', rng.normal(size=(8, 3)) * rng.uniform(0.1, 5))
    for rep in range(12):
        p = rng.normal(size=(8, 3))
        p[rng.integers(0, 8, size=3)] = p[0]                   # repeated rows
        add('repeated rows', p)
        q = rng.normal(size=(8, 3))
        q[:, rng.integers(0, 3)] = np.round(q[:, rng.integers(0, 3)])      # repeated coordinates along one axis
        add('repeated coordinates', q)
    out['n'] = np.array(k)
    path = os.path.join(ROOT, 'tests', 'golden', 'sort_bbox.npz')
    np.savez_compressed(path, **out)
    print(f'wrote {path}: {k} cases')


if __name__ == '__main__':
    main()
