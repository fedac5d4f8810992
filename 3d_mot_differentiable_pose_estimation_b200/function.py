"""Batched pose fit as a torch.autograd.Function over the C-ABI CUDA library.

`PoseFit.apply(noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt, ref_compat)` is the
batched operator SURVEY.md section 8b asks for; `pose_fit(...)` is its keyword-friendly face and
`pose_fit_raw(...)` returns the float64 records exactly as the kernels write them.

What one call replaces in the reference (per object, all of it on the CPU there):
  run_pose's zero padding + `backproject` + NOC gather (PoseEst/pose_estimation.py:256-290, :323),
  then `estimateSimilarityUmeyama` (PoseEst/pose_utils.py:16-61) when `sample_idx is None`, or
  `estimateSimilarityTransform` (pose_utils.py:86-117) with `np.random.randint` (:73) replaced by
  the host-supplied `sample_idx[b, h, :]`.
The backward pass has no reference counterpart (the upstream code detaches first,
Detection/tracker/postprocess.py:151).

PyTorch is plumbing here: it owns device memory and the stream; every computation happens in
libposefit_b200.so.  Nothing falls back to torch or the CPU: on a non-CUDA tensor the call raises.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from . import _lib

STATUS_OK, STATUS_EMPTY, STATUS_LOW_INLIER_RATIO, STATUS_NAN = 0, 1, 2, 3
# flag bits of `ref_compat` (include/posefit.h): True / 1 = the reference's scoring quirks; | SAMPLES_ARE_BITS = sample_idx
# holds uniform 32-bit draws that the kernel maps to floor(u * N / 2^32) (see device_sample_bits)
REF_COMPAT, SAMPLES_ARE_BITS = 1, 2


class PoseFitRaw(NamedTuple):
    pose: torch.Tensor          # [B,16] f64: s, R(9, true rotation), t(3), n_fit, ratio, pass_t
    ctx: torch.Tensor           # [B,32] f64 saved state for backward
    status: torch.Tensor        # [B] i32
    n_valid: torch.Tensor       # [B] i32
    inlier_mask: Optional[torch.Tensor]   # [B,H,W] u8 (RANSAC only)
    winner: Optional[torch.Tensor]        # [B] i32 (RANSAC only)


def ransac_iterations(raw: "PoseFitRaw") -> torch.Tensor:
    """[B] int64: how many iterations of getRANSACInliers' loop (PoseEst/pose_utils.py:72-82) the reference would have
    run for each object -- up to and including the one whose residual falls below StopT (:80-81), n_hyp when none does,
    0 for an object without correspondences.  Each iteration is one `np.random.randint(N, size=10)` (:73): this is what
    the drop-ins use to leave the global stream where the reference leaves it."""
    return raw.ctx[:, 30].to(torch.int64)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(lib, dev, b: int, h: int, w: int, n_hyp: int = 0, n_samp: int = 0) -> torch.Tensor:
    """Caller-owned scratch (partial moments / per-object records), sized by the library."""
    nbytes = int(lib.posefit_workspace_bytes(b, h, w, n_hyp, n_samp))
    return torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)


_KINV_CACHE = {}


def default_kinv(device=None, height: int = 240, width: int = 320) -> torch.Tensor:
    """K^-1 of the fixed MOTFront camera run_pose builds (pose_estimation.py:269-288), float64.
    Cached per (device, height, width): after the first call there is no host work and no host-to-device copy,
    so `kinv=None` (the documented default of every entry here) is asynchronous and CUDA-graph capturable."""
    dev = torch.device(device) if device is not None else torch.device('cpu')
    if dev.type == 'cuda' and dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    key = (dev.type, dev.index, int(height), int(width))
    k = _KINV_CACHE.get(key)
    if k is None:
        from .synth import motfront_intrinsics
        k = torch.linalg.inv(motfront_intrinsics(height, width)).contiguous()   # inv may return column-major strides
        k = k.to(dev)
        _KINV_CACHE[key] = k
    return k


def _prep_kinv(kinv, device, n_objects: int):
    if kinv is None:
        return default_kinv(device), 0
    if not (isinstance(kinv, torch.Tensor) and kinv.device == device and kinv.dtype == torch.float64
            and kinv.is_contiguous()):
        kinv = torch.as_tensor(kinv).to(device=device, dtype=torch.float64).contiguous()
    if kinv.shape == (3, 3):
        return kinv, 0
    if kinv.shape == (n_objects, 3, 3):
        return kinv, 1
    raise ValueError(f'kinv must be [3,3] or [B,3,3], got {tuple(kinv.shape)}')


def _check_crops(noc, depth, mask, bbox_xy0):
    if not noc.is_cuda:
        raise _lib.PoseFitError('pose_fit needs CUDA tensors: the solver has no CPU path')
    b, c, h, w = noc.shape
    if c != 3:
        raise ValueError('noc must be [B,3,H,W]')
    if depth.shape != (b, h, w) or mask.shape != (b, h, w) or bbox_xy0.shape != (b, 2):
        raise ValueError('depth/mask must be [B,H,W] and bbox_xy0 [B,2]')
    noc = noc.detach().to(torch.float32).contiguous()
    depth = depth.detach().to(device=noc.device, dtype=torch.float32).contiguous()
    if mask.dtype == torch.bool:
        mask = mask.to(torch.uint8)
    mask = mask.to(device=noc.device, dtype=torch.uint8).contiguous()
    bbox_xy0 = bbox_xy0.to(device=noc.device, dtype=torch.int32).contiguous()
    return noc, depth, mask, bbox_xy0, b, h, w


def pose_fit_raw(noc, depth, mask, bbox_xy0, kinv=None, sample_idx=None, ratio_adapt: float = 1.0,
                 ref_compat: bool = True, _f32=None, _valid_mask=None) -> PoseFitRaw:
    """Forward only, float64 records, no autograd.  Inputs: noc [B,3,H,W] f32 in [0,1],
    depth [B,H,W] f32, mask [B,H,W] u8/bool, bbox_xy0 [B,2] i32 (x0, y0 of each crop in the frame),
    kinv [3,3] or [B,3,3] f64 (None = MOTFront camera), sample_idx [B,n_hyp,n_samp] i32 or None.
    Nothing here touches the host when the tensors are already what the kernels take -- CUDA, contiguous, kinv float64
    (or None: cached per device), sample_idx int32 ON THE DEVICE; then the call is asynchronous and graph-capturable.
    A CPU `sample_idx` (or any tensor of another dtype) is converted and uploaded on every call: that is a host-to-device
    copy per call, not capturable -- draw on the host once per batch and pass the device tensor, or use
    `device_sample_bits` (draws made on the GPU).
    `_f32` = (scale[B], rot[B,9], trans[B,3]) float32 tensors and `_valid_mask` [B,H,W] u8 are filled by the kernels
    when given (the autograd operator's outputs; posefit_forward_ex / posefit_forward_ransac_ex)."""
    lib = _lib.lib()
    noc, depth, mask, bbox_xy0, b, h, w = _check_crops(noc, depth, mask, bbox_xy0)
    dev = noc.device
    kinv, per_obj = _prep_kinv(kinv, dev, b)
    pose = torch.empty(b, _lib.POSE_DOUBLES, dtype=torch.float64, device=dev)
    ctx = torch.empty(b, _lib.CTX_DOUBLES, dtype=torch.float64, device=dev)
    status = torch.empty(b, dtype=torch.int32, device=dev)
    n_valid = torch.empty(b, dtype=torch.int32, device=dev)
    fs, fr, ft = _f32 if _f32 is not None else (None, None, None)
    with torch.cuda.device(dev):
        if sample_idx is None:
            ws = _workspace(lib, dev, b, h, w)
            code = lib.posefit_forward_ex(_ptr(noc), _ptr(depth), _ptr(mask), _ptr(bbox_xy0), _ptr(kinv), per_obj,
                                          b, h, w, _ptr(pose), _ptr(ctx), _ptr(status), _ptr(n_valid), _ptr(fs),
                                          _ptr(fr), _ptr(ft), _ptr(_valid_mask), _ptr(ws), ws.numel(), _stream(dev))
            _lib.check(code, 'posefit_forward')
            return PoseFitRaw(pose, ctx, status, n_valid, None, None)
        if not (sample_idx.device == dev and sample_idx.dtype == torch.int32 and sample_idx.is_contiguous()):
            sample_idx = sample_idx.to(device=dev, dtype=torch.int32).contiguous()
        if sample_idx.dim() != 3 or sample_idx.shape[0] != b:
            raise ValueError('sample_idx must be [B,n_hyp,n_samp]')
        n_hyp, n_samp = int(sample_idx.shape[1]), int(sample_idx.shape[2])
        inl = torch.empty(b, h, w, dtype=torch.uint8, device=dev)
        winner = torch.empty(b, dtype=torch.int32, device=dev)
        ws = _workspace(lib, dev, b, h, w, max(n_hyp, 1), n_samp)
        code = lib.posefit_forward_ransac_ex(_ptr(noc), _ptr(depth), _ptr(mask), _ptr(bbox_xy0), _ptr(kinv), per_obj,
                                             _ptr(sample_idx), b, h, w, n_hyp, n_samp, float(ratio_adapt),
                                             int(ref_compat), _ptr(pose), _ptr(ctx), _ptr(status),
                                             _ptr(n_valid), _ptr(inl), _ptr(winner), _ptr(fs), _ptr(fr), _ptr(ft),
                                             _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(code, 'posefit_forward_ransac')
    return PoseFitRaw(pose, ctx, status, n_valid, inl, winner)


def device_sample_bits(n_objects: int, n_hyp: int, n_samp: int = 10, device='cuda', generator=None) -> torch.Tensor:
    """RANSAC draws made ON THE DEVICE (torch's Philox generator), before anyone knows how many correspondences an
    object has: uniform 32-bit values u, which the kernels map to floor(u * N / 2^32) when `ref_compat` carries
    SAMPLES_ARE_BITS.  Not numpy's stream -- host-supplied indices remain the parity mode (pose_utils.py:73)."""
    u = torch.randint(-2 ** 31, 2 ** 31, (n_objects, n_hyp, n_samp), dtype=torch.int64, device=device, generator=generator)
    return u.to(torch.int32)


def points_fit_raw(src, dst, mask=None, sample_idx=None, ratio_adapt: float = 1.0,
                   ref_compat: bool = True, pass_threshold: float = 0.0, stop_threshold: float = 0.0) -> PoseFitRaw:
    """Points mode: src/dst [B,3,N] float64 CUDA tensors (rows 0..2 of the reference's [4,N]
    homogeneous arrays), mask [B,N] u8 (None = all points)."""
    lib = _lib.lib()
    if not src.is_cuda:
        raise _lib.PoseFitError('points_fit needs CUDA tensors: the solver has no CPU path')
    b, c, n = src.shape
    if c != 3 or dst.shape != src.shape:
        raise ValueError('src and dst must both be [B,3,N]')
    dev = src.device
    src = src.detach().to(torch.float64).contiguous()
    dst = dst.detach().to(device=dev, dtype=torch.float64).contiguous()
    if mask is None:
        mask = torch.ones(b, n, dtype=torch.uint8, device=dev)
    mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
    pose = torch.empty(b, _lib.POSE_DOUBLES, dtype=torch.float64, device=dev)
    ctx = torch.empty(b, _lib.CTX_DOUBLES, dtype=torch.float64, device=dev)
    status = torch.empty(b, dtype=torch.int32, device=dev)
    n_valid = torch.empty(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        if sample_idx is None:
            ws = _workspace(lib, dev, b, 1, n)
            code = lib.posefit_points_forward(_ptr(src), _ptr(dst), _ptr(mask), b, n, _ptr(pose), _ptr(ctx),
                                              _ptr(status), _ptr(n_valid), _ptr(ws), ws.numel(), _stream(dev))
            _lib.check(code, 'posefit_points_forward')
            return PoseFitRaw(pose, ctx, status, n_valid, None, None)
        sample_idx = sample_idx.to(device=dev, dtype=torch.int32).contiguous()
        n_hyp, n_samp = int(sample_idx.shape[1]), int(sample_idx.shape[2])
        inl = torch.empty(b, n, dtype=torch.uint8, device=dev)
        winner = torch.empty(b, dtype=torch.int32, device=dev)
        ws = _workspace(lib, dev, b, 1, n, max(n_hyp, 1), n_samp)
        code = lib.posefit_points_forward_ransac(_ptr(src), _ptr(dst), _ptr(mask), _ptr(sample_idx), b, n, n_hyp,
                                                 n_samp, float(ratio_adapt), float(pass_threshold),
                                                 float(stop_threshold), int(bool(ref_compat)), _ptr(pose),
                                                 _ptr(ctx), _ptr(status), _ptr(n_valid), _ptr(inl), _ptr(winner),
                                                 _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(code, 'posefit_points_forward_ransac')
    return PoseFitRaw(pose, ctx, status, n_valid, inl, winner)


def pose_fit_backward_raw(noc, depth, mask, inlier_mask, bbox_xy0, kinv, ctx, status, grad_scale, grad_R, grad_t,
                          want_depth_grad: bool = False, out=None):
    """NOC (and optionally depth) gradient from the saved context.  All tensors on one CUDA device."""
    lib = _lib.lib()
    noc, depth, mask, bbox_xy0, b, h, w = _check_crops(noc, depth, mask, bbox_xy0)
    dev = noc.device
    kinv, per_obj = _prep_kinv(kinv, dev, b)

    def f32(t, shape):
        if t is None:
            return None
        return t.detach().to(device=dev, dtype=torch.float32).reshape(shape).contiguous()

    grad_scale, grad_R, grad_t = f32(grad_scale, (b,)), f32(grad_R, (b, 9)), f32(grad_t, (b, 3))
    g_noc = out if out is not None else torch.empty_like(noc)      # (out: a caller-owned [B,3,H,W] f32 buffer)
    g_depth = torch.empty_like(depth) if want_depth_grad else None
    if inlier_mask is not None:
        inlier_mask = inlier_mask.to(torch.uint8).contiguous()
    ws = torch.empty(max(int(lib.posefit_backward_workspace_bytes(b)), 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_backward(_ptr(noc), _ptr(depth), _ptr(mask), _ptr(inlier_mask), _ptr(bbox_xy0), _ptr(kinv),
                                    per_obj, b, h, w, _ptr(ctx), _ptr(status), _ptr(grad_scale), _ptr(grad_R),
                                    _ptr(grad_t), _ptr(g_noc), _ptr(g_depth), _ptr(ws), ws.numel(), _stream(dev))
    _lib.check(code, 'posefit_backward')
    return g_noc, g_depth


class PoseEpilogue(NamedTuple):
    global_rot: torch.Tensor     # [B,3,3] f64, scale embedded (pose_estimation.py:404-406)
    global_trans: torch.Tensor   # [B,3]
    global_scale: torch.Tensor   # [B]
    euler: torch.Tensor          # [B,3] XYZ Euler angles of the unscaled rotation (postprocess.py:158-160)
    world_box: torch.Tensor      # [B,8,3] sort_bbox-ordered world box of the depth points (:373-380)


def pose_epilogue(raw: PoseFitRaw, depth, mask, bbox_xy0, kinv=None, campose=None, cam_index=None) -> PoseEpilogue:
    """Batched tail of run_pose on the GPU (no per-instance Python loop, no host round trips).
    campose: [4,4], [F,4,4] with cam_index[B], or [B,4,4] camera-to-world matrices; None keeps camera
    space (run_pose_office)."""
    lib = _lib.lib()
    dev = raw.pose.device
    b = int(raw.pose.shape[0])
    depth = depth.detach().to(device=dev, dtype=torch.float32).contiguous()
    h, w = int(depth.shape[1]), int(depth.shape[2])
    mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
    bbox_xy0 = bbox_xy0.to(device=dev, dtype=torch.int32).contiguous()
    kinv, per_obj = _prep_kinv(kinv, dev, b)
    n_cam = 0
    if campose is not None:
        campose = torch.as_tensor(campose).to(device=dev, dtype=torch.float64).reshape(-1, 16).contiguous()
        n_cam = int(campose.shape[0])
    if cam_index is not None:
        cam_index = cam_index.to(device=dev, dtype=torch.int32).contiguous()
    elif n_cam not in (0, 1, b):
        raise ValueError('campose must be [4,4] or [B,4,4] unless cam_index is given')
    out = torch.empty(b, 40, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_epilogue(_ptr(depth), _ptr(mask), _ptr(bbox_xy0), _ptr(kinv), per_obj, _ptr(raw.pose),
                                    _ptr(raw.status), _ptr(campose), n_cam, _ptr(cam_index), b, h, w, _ptr(out),
                                    _stream(dev))
    _lib.check(code, 'posefit_epilogue')
    return PoseEpilogue(out[:, 0:9].reshape(b, 3, 3), out[:, 9:12], out[:, 12], out[:, 13:16],
                        out[:, 16:40].reshape(b, 8, 3))


def clip_mask_to_box(depth, mask, bbox_xy0, gt_box, campose, kinv=None, cam_index=None, min_keep: int = 20):
    """GT-box pre-filter of run_pose (clean_depth, pose_estimation.py:107-134, :293-299) as a mask:
    returns (new_mask [B,H,W] u8, kept [B] i32).  gt_box [B,8,3]; campose [4,4], [B,4,4] or [F,4,4] +
    cam_index[B]."""
    lib = _lib.lib()
    if not depth.is_cuda:
        raise _lib.PoseFitError('clip_mask_to_box needs CUDA tensors: the solver has no CPU path')
    dev = depth.device
    b, h, w = (int(v) for v in depth.shape)
    depth = depth.detach().to(torch.float32).contiguous()
    mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
    bbox_xy0 = bbox_xy0.to(device=dev, dtype=torch.int32).contiguous()
    kinv, per_obj = _prep_kinv(kinv, dev, b)
    gt_box = torch.as_tensor(gt_box).to(device=dev, dtype=torch.float64).reshape(b, 24).contiguous()
    campose = torch.as_tensor(campose).to(device=dev, dtype=torch.float64).reshape(-1, 16).contiguous()
    n_cam = int(campose.shape[0])
    if cam_index is not None:
        cam_index = cam_index.to(device=dev, dtype=torch.int32).contiguous()
    elif n_cam not in (1, b):
        raise ValueError('campose must be [4,4] or [B,4,4] unless cam_index is given')
    out = torch.empty_like(mask)
    kept = torch.empty(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_clip_mask(_ptr(depth), _ptr(mask), _ptr(bbox_xy0), _ptr(kinv), per_obj, _ptr(campose), n_cam,
                                     _ptr(cam_index), _ptr(gt_box), int(min_keep), b, h, w, _ptr(out), _ptr(kept),
                                     _stream(dev))
    _lib.check(code, 'posefit_clip_mask')
    return out, kept


def statistical_outlier_mask(noc, depth, mask, bbox_xy0, kinv=None, source: str = 'depth', nb_neighbors: int = 20,
                             std_ratio: float = 2.0, min_points: int = 100):
    """Open3D-style statistical outlier removal as a mask filter (run_pose, pose_estimation.py:311-318
    for source='depth', :341-349 for source='noc').  UNPINNED semantics (Open3D is not vendored).
    Returns the surviving subset of mask & depth>0 as [B,H,W] u8."""
    lib = _lib.lib()
    if not depth.is_cuda:
        raise _lib.PoseFitError('statistical_outlier_mask needs CUDA tensors: the solver has no CPU path')
    dev = depth.device
    b, h, w = (int(v) for v in depth.shape)
    depth = depth.detach().to(torch.float32).contiguous()
    mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
    bbox_xy0 = bbox_xy0.to(device=dev, dtype=torch.int32).contiguous()
    if noc is not None:
        noc = noc.detach().to(device=dev, dtype=torch.float32).contiguous()
    kinv, per_obj = _prep_kinv(kinv, dev, b)
    ws = torch.empty(max(int(lib.posefit_sor_workspace_bytes(b, h, w)), 8), dtype=torch.uint8, device=dev)
    out = torch.empty_like(mask)
    with torch.cuda.device(dev):
        code = lib.posefit_sor_mask(_ptr(noc), _ptr(depth), _ptr(mask), _ptr(bbox_xy0), _ptr(kinv), per_obj,
                                    {'depth': 0, 'noc': 1}[source], int(nb_neighbors), float(std_ratio),
                                    int(min_points), b, h, w, _ptr(out), _ptr(ws), ws.numel(), _stream(dev))
    _lib.check(code, 'posefit_sor_mask')
    return out


def _forward_outputs(ctx, noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt, ref_compat, return_mask):
    """Shared forward of PoseFit / PoseFitFull.  Everything the operator returns is written by the library's kernels:
    scale / R / t as float32 by the solve kernel, the plain fit's "inlier" mask (every valid correspondence,
    pose_estimation.py:23-25) by the moments kernel -- no eager torch arithmetic on the batch."""
    b, _, h, w = noc.shape
    dev = noc.device
    if not noc.is_cuda:
        raise _lib.PoseFitError('pose_fit needs CUDA tensors: the solver has no CPU path')
    scale = torch.empty(b, dtype=torch.float32, device=dev)
    rot = torch.empty(b, 3, 3, dtype=torch.float32, device=dev)
    trans = torch.empty(b, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(b, h, w, dtype=torch.uint8, device=dev) if (sample_idx is None and return_mask) else None
    raw = pose_fit_raw(noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt, ref_compat, _f32=(scale, rot, trans),
                       _valid_mask=valid)
    inl = raw.inlier_mask if raw.inlier_mask is not None else valid
    ctx.has_inliers = raw.inlier_mask is not None
    ctx.kinv = kinv
    ctx.depth_grad = bool(depth.requires_grad)
    ctx.in_dtypes = (noc.dtype, depth.dtype)
    ctx.save_for_backward(noc, depth, mask, bbox_xy0, raw.ctx, raw.status,
                          raw.inlier_mask if raw.inlier_mask is not None else torch.empty(0, dtype=torch.uint8, device=dev))
    out_dtype = noc.dtype if noc.dtype.is_floating_point else torch.float32
    if out_dtype != torch.float32:
        scale, rot, trans = scale.to(out_dtype), rot.to(out_dtype), trans.to(out_dtype)
    return scale, rot, trans, inl, raw


class PoseFit(torch.autograd.Function):
    """(scale[B], R[B,3,3], t[B,3], inlier_mask[B,H,W] u8, status[B] i32, n_valid[B] i32) =
    PoseFit.apply(noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt, ref_compat, return_mask).

    inlier_mask: the RANSAC winner's inlier set.  For the plain fit (sample_idx=None) every valid correspondence takes
    part (mask != 0 and depth > 0, pose_estimation.py:23-25); that mask is only materialised -- by the moments kernel,
    1 B/pixel of extra traffic -- when return_mask=True, otherwise the slot is None.

    R is the true rotation (the reference's `Rotation` is R^T, pose_utils.py:44), so the
    object-to-camera matrix of run_pose (pose_estimation.py:401-403) is [s*R | t].  Outputs have
    noc's dtype (float32); `pose_fit_raw` gives the float64 records.  Gradients flow to `noc`
    and, if it requires grad, to `depth`; the RANSAC selection is a constant."""

    @staticmethod
    def forward(ctx, noc, depth, mask, bbox_xy0, kinv=None, sample_idx=None, ratio_adapt=1.0, ref_compat=True,
                return_mask=False):
        scale, rot, trans, inl, raw = _forward_outputs(ctx, noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt,
                                                       ref_compat, return_mask)
        ctx.mark_non_differentiable(*[t for t in (inl, raw.status, raw.n_valid) if t is not None])
        return scale, rot, trans, inl, raw.status, raw.n_valid

    @staticmethod
    def backward(ctx, g_scale, g_rot, g_trans, *_unused):
        noc, depth, mask, bbox_xy0, saved, status, inl = ctx.saved_tensors
        g_noc, g_depth = pose_fit_backward_raw(noc, depth, mask, inl if ctx.has_inliers else None, bbox_xy0, ctx.kinv,
                                               saved, status, g_scale, g_rot, g_trans, ctx.depth_grad)
        g_noc = g_noc.to(ctx.in_dtypes[0])
        if g_depth is not None:
            g_depth = g_depth.to(ctx.in_dtypes[1])
        return g_noc, g_depth, None, None, None, None, None, None, None


def pose_fit(noc, depth, mask, bbox_xy0, kinv=None, sample_idx=None, ratio_adapt: float = 1.0,
             ref_compat: bool = True, return_mask: bool = False):
    """Keyword-friendly PoseFit.apply."""
    return PoseFit.apply(noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt, ref_compat, return_mask)


class PoseFitHead(torch.autograd.Function):
    """(scale[B], R[B,3,3], t[B,3], status[B] i32, n_valid[B] i32) =
    PoseFitHead.apply(noc_head, roi_hw, depth, mask, bbox_xy0, kinv)

    The plain fit fed by the NOC head output: `noc_head` [B,3,Hh,Wh] (the 3x28x28 sigmoid maps of the NOC head,
    Detection/roi_heads/nocs_head.py:232-235) is resized to every instance's box `roi_hw[b] = (h_b, w_b)` ON THE FLY inside
    the fit's loaders (posefit_forward_head) -- the roi_align resize of postprocess.py:141-147 is never materialised --
    and the backward pass returns the gradient with respect to the head output (and the depth crop when it requires
    grad).  Same result as `pose_fit(resample_noc(noc_head, roi_hw, H, W), depth, mask, bbox_xy0, kinv)`, with 12 B/pixel
    less traffic in each of the four passes that composition makes."""

    @staticmethod
    def forward(ctx, noc_head, roi_hw, depth, mask, bbox_xy0, kinv=None):
        lib = _lib.lib()
        if not noc_head.is_cuda:
            raise _lib.PoseFitError('pose_fit_head needs CUDA tensors: the solver has no CPU path')
        dev = noc_head.device
        head = noc_head.detach().to(torch.float32).contiguous()
        b, c, hh, wh = head.shape
        if c != 3:
            raise ValueError('noc_head must be [B,3,Hh,Wh]')
        h, w = int(depth.shape[1]), int(depth.shape[2])
        if depth.shape != (b, h, w) or mask.shape != (b, h, w) or bbox_xy0.shape != (b, 2) or roi_hw.shape != (b, 2):
            raise ValueError('depth/mask must be [B,H,W], bbox_xy0 and roi_hw [B,2]')
        depth_c = depth.detach().to(device=dev, dtype=torch.float32).contiguous()
        mask_c = mask.to(device=dev, dtype=torch.uint8).contiguous()
        xy0 = bbox_xy0.to(device=dev, dtype=torch.int32).contiguous()
        roi = roi_hw.to(device=dev, dtype=torch.int32).contiguous()
        k, per_obj = _prep_kinv(kinv, dev, b)
        pose = torch.empty(b, _lib.POSE_DOUBLES, dtype=torch.float64, device=dev)
        saved = torch.empty(b, _lib.CTX_DOUBLES, dtype=torch.float64, device=dev)
        status = torch.empty(b, dtype=torch.int32, device=dev)
        n_valid = torch.empty(b, dtype=torch.int32, device=dev)
        scale = torch.empty(b, dtype=torch.float32, device=dev)
        rot = torch.empty(b, 3, 3, dtype=torch.float32, device=dev)
        trans = torch.empty(b, 3, dtype=torch.float32, device=dev)
        ws = torch.empty(max(int(lib.posefit_head_workspace_bytes(b)), 8), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            code = lib.posefit_forward_head(_ptr(head), _ptr(roi), _ptr(depth_c), _ptr(mask_c), _ptr(xy0), _ptr(k), per_obj,
                                            b, hh, wh, h, w, _ptr(pose), _ptr(saved), _ptr(status), _ptr(n_valid),
                                            _ptr(scale), _ptr(rot), _ptr(trans), _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(code, 'posefit_forward_head')
        ctx.kinv = kinv
        ctx.depth_grad = bool(depth.requires_grad)
        ctx.in_dtypes = (noc_head.dtype, depth.dtype)
        ctx.save_for_backward(head, roi, depth_c, mask_c, xy0, saved, status)
        ctx.mark_non_differentiable(status, n_valid)
        ctx.pose64 = pose
        return scale, rot, trans, status, n_valid

    @staticmethod
    def backward(ctx, g_scale, g_rot, g_trans, *_unused):
        lib = _lib.lib()
        head, roi, depth, mask, xy0, saved, status = ctx.saved_tensors
        dev = head.device
        b, _, hh, wh = head.shape
        h, w = int(depth.shape[1]), int(depth.shape[2])
        k, per_obj = _prep_kinv(ctx.kinv, dev, b)

        def f32(t, shape):
            return None if t is None else t.detach().to(device=dev, dtype=torch.float32).reshape(shape).contiguous()
        g_scale, g_rot, g_trans = f32(g_scale, (b,)), f32(g_rot, (b, 9)), f32(g_trans, (b, 3))
        g_head = torch.empty_like(head)
        g_depth = torch.empty_like(depth) if ctx.depth_grad else None
        ws = torch.empty(max(int(lib.posefit_backward_workspace_bytes(b)), 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            code = lib.posefit_backward_head(_ptr(head), _ptr(roi), _ptr(depth), _ptr(mask), None, _ptr(xy0), _ptr(k),
                                             per_obj, b, hh, wh, h, w, _ptr(saved), _ptr(status), _ptr(g_scale),
                                             _ptr(g_rot), _ptr(g_trans), _ptr(g_head), _ptr(g_depth), _ptr(ws), ws.numel(),
                                             _stream(dev))
        _lib.check(code, 'posefit_backward_head')
        g_head = g_head.to(ctx.in_dtypes[0])
        if g_depth is not None:
            g_depth = g_depth.to(ctx.in_dtypes[1])
        return g_head, None, g_depth, None, None, None


def pose_fit_head(noc_head, roi_hw, depth, mask, bbox_xy0, kinv=None):
    """Keyword-friendly PoseFitHead.apply."""
    return PoseFitHead.apply(noc_head, roi_hw, depth, mask, bbox_xy0, kinv)


class PoseFitFull(torch.autograd.Function):
    """PoseFit that also hands out the float64 pose records and the RANSAC winners (non-differentiable), for
    callers that feed both autograd (scale, R, t) and the batched epilogue (records): one forward, not two.
    Returns (scale, R, t, inlier_mask, status, n_valid, pose64 [B,16], winner [B] or empty, ctx64 [B,32]);
    ctx64[:, 30] = RANSAC iterations the reference's loop would have run (`ransac_iterations`)."""

    @staticmethod
    def forward(ctx, noc, depth, mask, bbox_xy0, kinv=None, sample_idx=None, ratio_adapt=1.0, ref_compat=True,
                return_mask=False):
        scale, rot, trans, inl, raw = _forward_outputs(ctx, noc, depth, mask, bbox_xy0, kinv, sample_idx, ratio_adapt,
                                                       ref_compat, return_mask)
        winner = raw.winner if raw.winner is not None else torch.empty(0, dtype=torch.int32, device=noc.device)
        ctx.mark_non_differentiable(*[t for t in (inl, raw.status, raw.n_valid, raw.pose, winner, raw.ctx) if t is not None])
        return scale, rot, trans, inl, raw.status, raw.n_valid, raw.pose, winner, raw.ctx

    @staticmethod
    def backward(ctx, g_scale, g_rot, g_trans, *_unused):
        return PoseFit.backward(ctx, g_scale, g_rot, g_trans)
