"""CPU oracle for the PoseEst hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy float64 restatement of the reference's per-object 7-DoF pose solver
(`PoseEst/pose_utils.py` + the solver part of `PoseEst/pose_estimation.py` of
DomiSchmauser/3D_MOT_Differentiable_Pose_Estimation).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the shipped package never does.

Parity pinning: the reference ships no tests or golden vectors,
so this restatement is pinned against outputs of the *real* reference modules,
imported from /root/reference by `oracle/gen_golden.py` (see `oracle/ref_import.py`)
and committed under `tests/golden/`.  `tests/test_oracle_golden.py` checks every
function below against those fixtures.  Two things stay "parity unpinned":
  * gradients (the reference never back-propagates through the fit)
    -- see `oracle/grad_oracle.py`;
  * the Open3D statistical-outlier filters and the GT-box clip inside `run_pose`
    (third-party open3d==0.10.0.0 is not vendored), which are outside the path.

Conventions.  Point sets are `[N, 3]` float64 (the reference uses homogeneous
`[4, N]`; the arithmetic is the same).  "rot_t" is what the reference calls
`Rotation`: the TRANSPOSE of the true rotation R (pose_utils.py:44).  All
reference quirks are reproduced on purpose.
"""
from __future__ import annotations

import numpy as np

# fixed MOTFront camera, pose_estimation.py:269-288
FRAME_H, FRAME_W = 240, 320
FOCAL = 292.87803547399


def motfront_intrinsics(height: int = FRAME_H, width: int = FRAME_W) -> np.ndarray:
    """K as built in run_pose (pose_estimation.py:269-288)."""
    return np.array([[FOCAL, 0.0, width / 2 - 0.5],
                     [0.0, FOCAL, height / 2 - 0.5],
                     [0.0, 0.0, 1.0]])


# --------------------------------------------------------------------------- #
# back-projection  (pose_estimation.py:16-43)
# --------------------------------------------------------------------------- #
def backproject_points(depth: np.ndarray, intrinsics: np.ndarray, bin_mask: np.ndarray):
    """Camera-space points of all pixels with `mask & depth>0`, row-major order.

    Follows pose_estimation.py:22-41: K^-1 * [u, v, 1], scaled so that the third
    component equals the depth, then y and z negated.  Returns (pts[N,3], (rows, cols)).
    """
    k_inv = np.linalg.inv(intrinsics)                       # :22
    keep = np.logical_and(bin_mask, depth > 0)              # :23-25
    rows, cols = np.where(keep)                             # :27  (row-major)
    uv1 = np.stack([cols, rows, np.ones(rows.shape[0])])    # :28-32  [3, N]
    rays = (k_inv @ uv1).T                                  # :34-35
    z = depth[rows, cols]                                   # :37
    pts = rays * z[:, None] / rays[:, 2:3]                  # :39
    pts[:, 1] = -pts[:, 1]                                  # :40
    pts[:, 2] = -pts[:, 2]                                  # :41
    return pts, (rows, cols)


# --------------------------------------------------------------------------- #
# Umeyama / Procrustes fit  (pose_utils.py:16-61)
# --------------------------------------------------------------------------- #
def umeyama_fit(src: np.ndarray, dst: np.ndarray):
    """Similarity fit dst ~ s*R*src + t on `[N,3]` sets.

    Returns (scales[3], rot_t[3,3], trans[3], out_transform[4,4]) exactly as the
    reference does: rot_t = (U Vh)^T = R^T (:44), trans = mu_dst - s*R*mu_src (:55),
    out_transform[:3,:3] = s * rot_t (:58) -- the transposed block is the reference's
    own quirk and is what RANSAC scores with.
    """
    n = src.shape[0]
    mu_s = np.mean(src.T, axis=1)                           # :23
    mu_d = np.mean(dst.T, axis=1)                           # :24
    cs = src.T - mu_s[:, None]                              # :27
    cd = dst.T - mu_d[:, None]                              # :28
    cov = (cd @ cs.T) / n                                   # :30
    if np.isnan(cov).any():                                 # :32-36
        raise RuntimeError('There are NANs in the input.')
    u, d, vh = np.linalg.svd(cov, full_matrices=True)       # :38
    if np.linalg.det(u) * np.linalg.det(vh) < 0.0:          # :39-42
        d[-1] = -d[-1]
        u[:, -1] = -u[:, -1]
    rot_t = (u @ vh).T                                      # :44
    var_s = np.var(src.T, axis=1).sum()                     # :46 (ddof 0)
    if var_s * np.sum(d) != 0:                              # :47-50
        s = 1 / var_s * np.sum(d)
    else:
        s = 1
    scales = np.array([s, s, s])
    trans = dst.T.mean(axis=1) - src.T.mean(axis=1).dot(s * rot_t)   # :55
    out = np.identity(4)
    out[:3, :3] = np.diag(scales) @ rot_t                   # :58
    out[:3, 3] = trans                                      # :59
    return scales, rot_t, trans, out


# --------------------------------------------------------------------------- #
# hypothesis scoring  (pose_utils.py:5-14)
# --------------------------------------------------------------------------- #
def score_model(out_transform: np.ndarray, src: np.ndarray, dst: np.ndarray, pass_t: float):
    """(total residual, inlier ratio with the index-0 quirk, inlier indices, per-point r)."""
    diff = dst.T - (out_transform[:3, :3] @ src.T + out_transform[:3, 3:4])   # :7
    r = np.linalg.norm(diff, axis=0)                        # :8
    residual = np.linalg.norm(r)                            # :9
    inl = np.where(r < pass_t)                              # :10
    n_counted = np.count_nonzero(inl)                       # :11  counts non-zero INDEX values (F5)
    return residual, n_counted / src.shape[0], inl[0], r


# --------------------------------------------------------------------------- #
# RANSAC with host-supplied sample indices  (pose_utils.py:63-83)
# --------------------------------------------------------------------------- #
def ransac_inliers(src, dst, sample_idx, pass_t, stop_t):
    """Replays getRANSACInliers with `sample_idx[n_hyp, n_samp]` instead of np.random.randint.

    Returns a dict with the winning hypothesis, its inlier indices, the (quirky) inlier
    ratio, per-hypothesis residuals (NaN for hypotheses after the early stop) and the
    smallest relative distance of any per-point residual of the WINNER to the pass
    threshold (the "margin" the parity tests report).
    """
    n = src.shape[0]
    best_res, best_ratio, best_idx, best_h = 1e10, 0, np.arange(n), -1      # :68-70
    best_r = None
    residuals = np.full(sample_idx.shape[0], np.nan)
    for h in range(sample_idx.shape[0]):                    # :71
        pick = sample_idx[h]                                # :73 (replayed)
        _, _, _, tf = umeyama_fit(src[pick], dst[pick])     # :74
        res, ratio, idx, r = score_model(tf, src, dst, pass_t)   # :75
        residuals[h] = res
        if res < best_res:                                  # :76-79
            best_res, best_ratio, best_idx, best_h, best_r = res, ratio, idx, h, r
        if best_res < stop_t:                               # :80-81
            break
    margin = np.inf
    if best_r is not None and pass_t > 0 and np.isfinite(pass_t):
        margin = float(np.min(np.abs(best_r - pass_t)) / pass_t)
    # iterations the loop ran (each one np.random.randint(N, size=10) in the reference, :73): residuals stay NaN
    # behind the early stop
    iterations = int(np.count_nonzero(~np.isnan(residuals))) if np.isnan(residuals).any() else int(sample_idx.shape[0])
    return dict(inlier_idx=best_idx, ratio=best_ratio, winner=best_h, residuals=residuals,
                best_residual=best_res, margin=margin, iterations=iterations)


def pass_thresholds(src, dst, ratio_adapt=1.0):
    """PassT / StopT heuristics, pose_utils.py:91-96."""
    t_norm = np.mean(np.linalg.norm(dst, axis=1))           # :91
    s_norm = np.mean(np.linalg.norm(src, axis=1))           # :92
    ts = t_norm / s_norm                                    # :93
    st = s_norm / t_norm                                    # :94
    pass_t = st * ratio_adapt if st > ts else ts * ratio_adapt   # :95
    return pass_t, pass_t / 100                             # :96


def similarity_transform(src, dst, sample_idx, ratio_adapt=1.0):
    """estimateSimilarityTransform (pose_utils.py:86-117) with replayed sample indices.

    Returns a dict: ok (False <=> the reference returns 4xNone, :105-107), scales, rot_t,
    trans, out_transform, inlier_idx, ratio, winner, margin, pass_t.
    """
    with np.errstate(divide='ignore', invalid='ignore'):
        pass_t, stop_t = pass_thresholds(src, dst, ratio_adapt)
    rr = ransac_inliers(src, dst, sample_idx, pass_t, stop_t)            # :103
    out = dict(ok=False, pass_t=pass_t, stop_t=stop_t, **rr)
    if rr['ratio'] < 0.1:                                   # :105-107
        return out
    keep = rr['inlier_idx']
    scales, rot_t, trans, tf = umeyama_fit(src[keep], dst[keep])         # :109
    out.update(ok=True, scales=scales, rot_t=rot_t, trans=trans, out_transform=tf)
    return out


# --------------------------------------------------------------------------- #
# per-object driver: run_pose minus the Open3D / GT-box filters
# (pose_estimation.py:256-290, 320-323, 359-367, 401-412)
# --------------------------------------------------------------------------- #
def crop_correspondences(noc_crop_hwc, depth_frame, mask_frame, abs_bbox, intrinsics=None):
    """NOC/depth correspondences of one instance, as run_pose builds them.

    noc_crop_hwc: [h, w, 3] in [0,1]; depth_frame / mask_frame: full [240, 320];
    abs_bbox: integer XYXY.  Returns (noc_pts[N,3] = noc-0.5, depth_pts[N,3], (rows, cols)).
    """
    x0, y0, x1, y1 = (int(v) for v in abs_bbox)
    fh, fw = depth_frame.shape
    depth_pad = np.zeros((fh, fw))                                              # :260
    depth_pad[y0:y1, x0:x1] = np.asarray(depth_frame, dtype=np.float32)[y0:y1, x0:x1]   # :261
    noc_pad = np.zeros((fh, fw, 3))                                             # :265
    noc_pad[y0:y1, x0:x1, :] = noc_crop_hwc                                     # :266
    if intrinsics is None:
        intrinsics = motfront_intrinsics(fh, fw)                                # :269-288
    depth_pts, (rows, cols) = backproject_points(depth_pad, intrinsics, np.asarray(mask_frame))   # :290
    noc_pts = noc_pad[rows, cols, :] - 0.5                                      # :323
    return noc_pts, depth_pts, (rows, cols)


def pose_from_correspondences(noc_pts, depth_pts, sample_idx=None, ratio_adapt=1.0):
    """Fit one object.  sample_idx=None -> plain Umeyama on all points (BASELINE config 2),
    else the RANSAC path.  Returns dict(status, s, R (true rotation), rot_t, t, inlier_idx, ...).

    status: 0 ok, 1 no correspondences (pose_estimation.py:361-362),
            2 inlier ratio < 0.1 (pose_utils.py:105-107), 3 NaN covariance (:32-36).
    """
    n = noc_pts.shape[0]
    if n == 0:
        return dict(status=1, n_valid=0)
    try:
        if sample_idx is None:
            scales, rot_t, trans, tf = umeyama_fit(noc_pts, depth_pts)
            res = dict(ok=True, scales=scales, rot_t=rot_t, trans=trans, out_transform=tf,
                       inlier_idx=np.arange(n), ratio=1.0, winner=-1, margin=np.inf)
        else:
            res = similarity_transform(noc_pts, depth_pts, sample_idx, ratio_adapt)
    except RuntimeError:
        return dict(status=3, n_valid=n)
    if not res['ok']:
        return dict(status=2, n_valid=n, **{k: res[k] for k in ('inlier_idx', 'ratio', 'winner', 'margin', 'pass_t', 'iterations')})
    out = dict(status=0, n_valid=n, s=float(res['scales'][0]), rot_t=res['rot_t'], R=res['rot_t'].T,
               t=res['trans'])
    for k in ('inlier_idx', 'ratio', 'winner', 'margin', 'pass_t', 'residuals', 'iterations'):
        if k in res:
            out[k] = res[k]
    return out


def object_to_camera(scales, rot_t, trans):
    """obj_tocam of run_pose (pose_estimation.py:401-403): [diag(S) * Rotation^T | t]."""
    m = np.identity(4)
    m[:3, :3] = np.diag(scales) @ rot_t.T
    m[:3, 3] = trans
    return m


def apply_similarity(scale, rot_t, trans, pc):
    """transform_pc (pose_estimation.py:45-57), including its float32 4x4 intermediate (:50)."""
    rt = np.zeros((4, 4), dtype=np.float32)
    rt[:3, :3] = np.diag(scale) @ rot_t.transpose()
    rt[:3, 3] = trans
    rt[3, 3] = 1
    return (rt[:3, :3] @ pc.transpose() + rt[:3, 3:]).transpose()


def camera_to_world(cam_pc, campose):
    """cam2world (pose_estimation.py:59-70)."""
    return (np.dot(campose[:3, :3], cam_pc.transpose()) + campose[:3, 3:]).transpose()


# --------------------------------------------------------------------------- #
# GT-box pre-filter  (clean_depth, pose_estimation.py:107-134, used at :293-299)
# --------------------------------------------------------------------------- #
def clip_to_box(depth_pts, gt_box, campose):
    """Indices of the camera-space points whose world position lies strictly inside the
    axis-aligned extent of the 8x3 GT box (pose_estimation.py:113-131)."""
    lo, hi = gt_box.min(axis=0), gt_box.max(axis=0)                     # :113-118
    world = camera_to_world(depth_pts.copy(), campose)                  # :120-122
    keep = [i for i, q in enumerate(world)                              # :126-130
            if q[0] > lo[0] and q[0] < hi[0] and q[1] > lo[1] and q[1] < hi[1] and q[2] > lo[2] and q[2] < hi[2]]
    return np.asarray(keep, dtype=np.int64)


def clipped_correspondence_indices(depth_pts, gt_box, campose, min_keep=20):
    """run_pose's use of clean_depth (:293-299): the clip is only taken when more than 20 points survive."""
    keep = clip_to_box(depth_pts, gt_box, campose)
    if len(keep) > min_keep:                                            # :295
        return keep
    return np.arange(depth_pts.shape[0])


# --------------------------------------------------------------------------- #
# statistical outlier removal -- UNPINNED (Open3D 0.10 is not vendored / installed)
# --------------------------------------------------------------------------- #
def statistical_outlier_indices(points, nb_neighbors=20, std_ratio=2.0):
    """Restatement of open3d.geometry.PointCloud.remove_statistical_outlier as run_pose calls it
    (pose_estimation.py:312, :342): mean distance of every point to its nb_neighbors nearest
    points (itself included), threshold = mean + std_ratio * sample std, keep 0 < avg < threshold.
    Returns the kept indices in ascending order.  Parity unpinned: written from Open3D's published
    algorithm, not checked against the library."""
    from scipy.spatial import cKDTree
    n = points.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    dist, _ = cKDTree(points).query(points, k=min(nb_neighbors, n))
    dist = dist.reshape(n, -1)
    avg = dist.mean(axis=1)
    valid = avg > 0
    mean = avg[valid].sum() / n
    std = np.sqrt(((avg[valid] - mean) ** 2).sum() / (n - 1))
    return np.where(valid & (avg < mean + std_ratio * std))[0]


def run_pose_filters(noc_pts, depth_pts, min_points=100):
    """The two filter passes of run_pose (pose_estimation.py:311-318, :341-349): returns the indices
    (into the input correspondences) that survive both."""
    keep = np.arange(depth_pts.shape[0])
    if depth_pts.shape[0] > min_points:                                 # :311
        keep = statistical_outlier_indices(depth_pts)
    if len(keep) > min_points:                                          # :341
        keep = keep[statistical_outlier_indices(noc_pts[keep])]
    return keep


# --------------------------------------------------------------------------- #
# batched convenience used by tests / bench (loops the per-object oracle)
# --------------------------------------------------------------------------- #
def batch_pose(noc, depth, mask, bbox_xy0, intrinsics=None, sample_idx=None, ratio_adapt=1.0,
               frame_hw=(FRAME_H, FRAME_W)):
    """Run the oracle on crop-layout inputs (the layout the CUDA path consumes).

    noc [B,3,H,W] f32, depth [B,H,W] f32, mask [B,H,W] u8/bool, bbox_xy0 [B,2] int (x0,y0),
    intrinsics K [3,3] or [B,3,3] (None -> the fixed MOTFront camera);
    every crop is pasted into a zero frame exactly as run_pose pads it, so the code path
    (full-frame np.where, K^-1, gather) is the reference's.  Returns a list of per-object dicts
    with an extra 'inlier_mask' [H,W] u8 in crop coordinates.
    """
    b, _, h, w = noc.shape
    fh, fw = frame_hw
    outs = []
    for i in range(b):
        x0, y0 = int(bbox_xy0[i, 0]), int(bbox_xy0[i, 1])
        depth_frame = np.zeros((fh, fw), dtype=np.float32)
        mask_frame = np.zeros((fh, fw), dtype=bool)
        depth_frame[y0:y0 + h, x0:x0 + w] = depth[i]
        mask_frame[y0:y0 + h, x0:x0 + w] = mask[i] != 0
        noc_hwc = np.transpose(noc[i], (1, 2, 0))
        intr = None
        if intrinsics is not None:
            intr = intrinsics if intrinsics.ndim == 2 else intrinsics[i]
        noc_pts, depth_pts, (rows, cols) = crop_correspondences(
            noc_hwc, depth_frame, mask_frame, (x0, y0, x0 + w, y0 + h), intr)
        idx = None if sample_idx is None else np.asarray(sample_idx[i])
        o = pose_from_correspondences(noc_pts, depth_pts, idx, ratio_adapt)
        im = np.zeros((h, w), dtype=np.uint8)
        if o['status'] in (0, 2) and 'inlier_idx' in o and rows.size:
            keep = o['inlier_idx']
            im[rows[keep] - y0, cols[keep] - x0] = 1
        o['inlier_mask'] = im
        outs.append(o)
    return outs
