"""Drop-in for the solver part of the reference's `PoseEst/pose_estimation.py`: `backproject`,
`transform_pc`, `cam2world`, `sort_bbox`, `run_pose`, `run_pose_office` with the reference's
signatures and return types, computed by the CUDA library.

The GT-box clip `clean_depth` (:107-134, :293-299) is reproduced (posefit_clip_mask) when
`gt_3d_box` is given.  The two Open3D `remove_statistical_outlier` passes (:311-318, :341-349) are
reproduced by posefit_sor_mask from Open3D's published algorithm; open3d==0.10.0.0 is not vendored,
so that filter's parity is UNPINNED; set APPLY_STATISTICAL_FILTER = False to skip it.  The world box is the axis-aligned box of the depth
points in Open3D's corner order followed by the reference's own `sort_bbox`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .pose_utils import rewind_to_reference_stream
from .function import (pose_fit_raw, default_kinv, clip_mask_to_box, statistical_outlier_mask, ransac_iterations,
                       _ptr, _stream)

__all__ = ['backproject', 'transform_pc', 'cam2world', 'sort_bbox', 'run_pose', 'run_pose_office']

FOCAL = 292.87803547399                      # pose_estimation.py:272-273
N_ITERATIONS, N_SAMPLES = 100, 10            # pose_utils.py:97, :73
# run_pose applies Open3D's statistical outlier removal twice (:311-318, :341-349); our restatement
# of it is unpinned, so it can be switched off to fit on all mask & depth>0 correspondences
APPLY_STATISTICAL_FILTER = True


def _device():
    if not torch.cuda.is_available():
        raise _lib.PoseFitError('pose_estimation needs a CUDA device: the solver has no CPU path')
    return torch.device('cuda', torch.cuda.current_device())


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _compact(noc, depth, mask, xy0, kinv):
    """One object through posefit_compact -> (src[N,3] or None, dst[N,3], rows[N], cols[N]) on the GPU."""
    lib = _lib.lib()
    dev = depth.device
    kinv = kinv.to(device=dev, dtype=torch.float64).contiguous()
    h, w = depth.shape[-2:]
    p = h * w
    dst = torch.empty(1, p, 3, dtype=torch.float64, device=dev)
    src = torch.empty(1, p, 3, dtype=torch.float64, device=dev) if noc is not None else None
    rows = torch.empty(1, p, dtype=torch.int32, device=dev)
    cols = torch.empty(1, p, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_compact(_ptr(noc), _ptr(depth), _ptr(mask), _ptr(xy0), _ptr(kinv), 0, 1, h, w,
                                   _ptr(src), _ptr(dst), _ptr(rows), _ptr(cols), _ptr(count), _stream(dev))
    _lib.check(code, 'posefit_compact')
    n = int(count[0])                                     # host sync: the reference API returns sized arrays
    return (src[0, :n] if src is not None else None), dst[0, :n], rows[0, :n], cols[0, :n]


def backproject(depth, intrinsics, bin_mask):
    """pose_estimation.py:16-43 -> (pts [N,3] float64, (rows, cols)) in np.where order."""
    dev = _device()
    d = torch.as_tensor(_np(depth).astype(np.float32)).to(dev).contiguous()
    m = torch.as_tensor(_np(bin_mask).astype(np.uint8)).to(dev).contiguous()
    kinv = torch.from_numpy(np.linalg.inv(_np(intrinsics).astype(np.float64))).to(dev).contiguous()   # :22
    xy0 = torch.zeros(1, 2, dtype=torch.int32, device=dev)
    _, dst, rows, cols = _compact(None, d, m, xy0, kinv)
    return dst.cpu().numpy(), (rows.cpu().numpy().astype(np.int64), cols.cpu().numpy().astype(np.int64))


def _affine(matrix34: np.ndarray, pts: np.ndarray) -> np.ndarray:
    lib = _lib.lib()
    dev = _device()
    pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float64))
    n = pts.shape[0]
    if n == 0:
        return pts.copy()
    m = torch.from_numpy(np.ascontiguousarray(matrix34.astype(np.float64))).to(dev)
    x = torch.from_numpy(pts).to(dev)
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        code = lib.posefit_transform_points(_ptr(m), 0, _ptr(x), _ptr(out), 1, n, _stream(dev))
    _lib.check(code, 'posefit_transform_points')
    return out.cpu().numpy()


def transform_pc(scale, rot, trans, pc):
    """pose_estimation.py:45-57, including its float32 4x4 intermediate (:50)."""
    rt = np.zeros((3, 4), dtype=np.float32)
    rt[:3, :3] = np.diag(scale) @ np.asarray(rot).transpose()
    rt[:3, 3] = trans
    return _affine(rt, pc)


def cam2world(cam_pc, campose):
    """pose_estimation.py:59-70."""
    campose = _np(campose).astype(np.float64)
    return _affine(campose[:3, :4], cam_pc)


def sort_bbox(bboxs):
    """pose_estimation.py:72-93 (host-side: an 8-point argsort, kept in NumPy so ties break identically)."""
    sort_y = np.flip(np.argsort(bboxs[:, 1]))
    y_sorted = bboxs[sort_y]
    sort_yx = np.concatenate((np.flip(np.argsort(y_sorted[0:4, 0])), np.flip(np.argsort(y_sorted[4:8, 0])) + 4), axis=None)
    yx_sorted = y_sorted[sort_yx]
    sort_zyx = np.concatenate((np.flip(np.argsort(yx_sorted[0:2, 2])), np.argsort(yx_sorted[2:4, 2]) + 2,
                               np.flip(np.argsort(yx_sorted[4:6, 2])) + 4, np.argsort(yx_sorted[6:8, 2]) + 6), axis=None)
    return yx_sorted[sort_zyx]


def _aabb_corners(pts: np.ndarray) -> np.ndarray:
    """Open3D AxisAlignedBoundingBox.get_box_points() corner order (unpinned, see module docstring)."""
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    e = hi - lo
    return np.array([lo, lo + [e[0], 0, 0], lo + [0, e[1], 0], lo + [0, 0, e[2]],
                     hi, hi - [e[0], 0, 0], hi - [0, e[1], 0], hi - [0, 0, e[2]]])


def _run(nocs, depth, kinv, campose, bin_mask, abs_bbox, use_depth_box, gt_3d_box=None):
    dev = _device()
    x0, y0, x1, y1 = (int(v) for v in _np(abs_bbox).reshape(-1)[:4])
    h, w = y1 - y0, x1 - x0
    noc = torch.as_tensor(nocs).detach().to(dev, torch.float32)
    if noc.shape[:2] != (h, w):
        raise ValueError(f'nocs patch {tuple(noc.shape)} does not match bbox {h}x{w}')
    noc = noc.permute(2, 0, 1).contiguous()[None]                       # HxWx3 -> [1,3,h,w]
    d_full = torch.as_tensor(_np(depth).astype(np.float32)).to(dev)
    m_full = torch.as_tensor(_np(bin_mask)).to(dev)
    d = d_full[y0:y1, x0:x1].contiguous()[None]                         # :260-262
    m = (m_full[y0:y1, x0:x1] != 0).to(torch.uint8).contiguous()[None]
    xy0 = torch.tensor([[x0, y0]], dtype=torch.int32, device=dev)
    kinv = kinv.to(dev)
    if gt_3d_box is not None and campose is not None:                   # clean_depth, :293-299
        m, _ = clip_mask_to_box(d, m, xy0, torch.as_tensor(_np(gt_3d_box)).reshape(1, 8, 3),
                                torch.as_tensor(_np(campose)), kinv)
    if APPLY_STATISTICAL_FILTER:
        m = statistical_outlier_mask(None, d, m, xy0, kinv, source='depth')     # :311-318
        m = statistical_outlier_mask(noc, d, m, xy0, kinv, source='noc')        # :341-349
    src, dst, rows, cols = _compact(noc[0], d[0], m[0], xy0, kinv)
    n = int(dst.shape[0])
    if n == 0:                                                          # :361-362
        return None, None, None, None, None, None
    rng_state = np.random.get_state()
    idx = np.random.randint(n, size=(N_ITERATIONS, N_SAMPLES))          # pose_utils.py:73
    raw = pose_fit_raw(noc, d, m, xy0, kinv, sample_idx=torch.from_numpy(idx.astype(np.int32))[None])
    rewind_to_reference_stream(rng_state, n, int(ransac_iterations(raw)[0]), N_ITERATIONS)   # :80-81 stops the draws too
    status = int(raw.status[0])
    pose = raw.pose[0].cpu().numpy()
    if status == 2:
        print('[ WARN ] - Something is wrong. Small BestInlierRatio: ', pose[14])
        return None, None, None, None, None, None                      # :365-366
    if status == 3:
        raise RuntimeError('There are NANs in the input.')
    s = pose[0]
    scales = np.array([s, s, s])
    rot_t = pose[1:10].reshape(3, 3).T
    trans = pose[10:13]
    noc_pts, depth_pts = src.cpu().numpy(), dst.cpu().numpy()
    cam_pc = transform_pc(scales, rot_t, trans, noc_pts)                # :367
    obj_tocam = np.identity(4)
    obj_tocam[:3, :3] = np.diag(scales) @ rot_t.T                       # :402
    obj_tocam[:3, 3] = trans
    if campose is not None:
        campose = _np(campose).astype(np.float64)
        world_pc = cam2world(cam_pc, campose)                           # :370
        depth_world = cam2world(depth_pts, campose)
        global_transform = campose @ obj_tocam                          # :404
    else:                                                               # run_pose_office: stay in camera space
        world_pc, depth_world, global_transform = cam_pc, depth_pts, obj_tocam
    box_src = depth_world if use_depth_box else world_pc                # :374-380
    world_box = sort_bbox(_aabb_corners(box_src))
    return global_transform[:3, :3], global_transform[:3, 3], scales[0], world_box, depth_world, world_pc


def run_pose(nocs, depth, campose, bin_mask, abs_bbox, vis_obj=False, gt_pc=None, gt_3d_box=None, use_depth_box=True):
    """pose_estimation.py:245-412 -> (global_rot, global_trans, global_scale, world_box, depth_world,
    world_pc) or 6 x None.  gt_3d_box (8x3, world space) applies the reference's clean_depth clip
    (:107-134, :293-299); vis_obj / gt_pc (visualisation) are accepted and ignored."""
    return _run(nocs, depth, default_kinv(), campose, bin_mask, abs_bbox, use_depth_box, gt_3d_box)


def run_pose_office(nocs, depth, cam_intrinsics, bin_mask, abs_bbox, vis_obj=False, gt_pc=None, gt_3d_box=None,
                    use_depth_box=True):
    """pose_estimation.py:415-512: per-frame intrinsics, results stay in camera space."""
    k = _np(torch.squeeze(torch.as_tensor(cam_intrinsics))).astype(np.float64)
    kinv = torch.from_numpy(np.linalg.inv(k))
    depth = _np(torch.squeeze(torch.as_tensor(_np(depth))))
    return _run(nocs, depth, kinv, None, bin_mask, abs_bbox, use_depth_box)
