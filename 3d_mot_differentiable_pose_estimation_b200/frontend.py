"""Batched front end: what `postprocess_dets` does per instance before it calls run_pose
(Detection/tracker/postprocess.py:131-152), for a whole batch of instances in two launches.

  crops = gather_crops(depth_frames, inst_masks, boxes_xyxy, frame_of, H, W)
  noc   = resample_noc(noc_head, crops.roi_hw, H, W)          # differentiable w.r.t. noc_head
  out   = pose_fit(noc, crops.depth, crops.mask, crops.bbox_xy0, ...)
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _lib
from .function import _ptr, _stream


class Crops(NamedTuple):
    depth: torch.Tensor      # [B,H,W] f32, zero outside each instance's box
    mask: torch.Tensor       # [B,H,W] u8
    bbox_xy0: torch.Tensor   # [B,2] i32
    roi_hw: torch.Tensor     # [B,2] i32 (h_i, w_i)


def gather_crops(depth_frames, inst_masks, boxes_xyxy, frame_of: Optional[torch.Tensor], height: int, width: int) -> Crops:
    """depth_frames [F,FH,FW] f32, inst_masks [B,FH,FW] bool/u8 (full-frame instance masks),
    boxes_xyxy [B,4] int (x0,y0,x1,y1; pose_estimation.py:260-262 slices [y0:y1, x0:x1])."""
    lib = _lib.lib()
    if not depth_frames.is_cuda:
        raise _lib.PoseFitError('gather_crops needs CUDA tensors: the solver has no CPU path')
    dev = depth_frames.device
    if depth_frames.dim() == 2:
        depth_frames = depth_frames[None]
    depth_frames = depth_frames.detach().to(torch.float32).contiguous()
    inst_masks = inst_masks.to(device=dev, dtype=torch.uint8).contiguous()
    boxes = boxes_xyxy.to(device=dev, dtype=torch.int32).contiguous()
    b = int(inst_masks.shape[0])
    fh, fw = int(depth_frames.shape[1]), int(depth_frames.shape[2])
    if frame_of is not None:
        frame_of = frame_of.to(device=dev, dtype=torch.int32).contiguous()
    depth = torch.empty(b, height, width, dtype=torch.float32, device=dev)
    mask = torch.empty(b, height, width, dtype=torch.uint8, device=dev)
    xy0 = torch.empty(b, 2, dtype=torch.int32, device=dev)
    roi_hw = torch.empty(b, 2, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_gather_crops(_ptr(depth_frames), _ptr(inst_masks), _ptr(frame_of), _ptr(boxes), b, fh, fw,
                                        height, width, _ptr(depth), _ptr(mask), _ptr(xy0), _ptr(roi_hw), _stream(dev))
    _lib.check(code, 'posefit_gather_crops')
    return Crops(depth, mask, xy0, roi_hw)


def pack_mask(mask) -> torch.Tensor:
    """Host side of the one-bit-per-pixel wire format: mask (bool / 0-1 array or CPU tensor of any shape) ->
    uint8 CPU tensor of ceil(numel / 8) bytes, pixel i = bit (i & 7) of byte (i >> 3) over the flattened array."""
    arr = mask.detach().cpu().numpy() if isinstance(mask, torch.Tensor) else np.asarray(mask)
    return torch.from_numpy(np.packbits(arr.reshape(-1) != 0, bitorder='little'))


def unpack_mask(bits: torch.Tensor, shape) -> torch.Tensor:
    """bits: uint8 CUDA tensor written by `pack_mask` (after its copy to the device) -> uint8 mask of `shape` (0 / 1),
    the layout every fit entry takes."""
    if not bits.is_cuda:
        raise _lib.PoseFitError('unpack_mask needs a CUDA tensor: the solver has no CPU path')
    shape = tuple(int(v) for v in shape)
    n = 1
    for v in shape:
        n *= v
    bits = bits.contiguous()
    if bits.dtype != torch.uint8 or bits.numel() * 8 < n:
        raise ValueError('bits must be uint8 with at least ceil(numel / 8) bytes')
    mask = torch.empty(shape, dtype=torch.uint8, device=bits.device)
    with torch.cuda.device(bits.device):
        code = _lib.lib().posefit_unpack_mask(_ptr(bits), n, _ptr(mask), _stream(bits.device))
    _lib.check(code, 'posefit_unpack_mask')
    return mask


class ResampleNoc(torch.autograd.Function):
    """noc[B,3,H,W] = per-instance ROI-align resize of noc_head[B,3,Hh,Wh] to roi_hw[b] = (h_b, w_b),
    zero padded (postprocess.py:141-147).  Gradient flows to noc_head."""

    @staticmethod
    def forward(ctx, noc_head, roi_hw, height, width):
        lib = _lib.lib()
        if not noc_head.is_cuda:
            raise _lib.PoseFitError('resample_noc needs CUDA tensors: the solver has no CPU path')
        dev = noc_head.device
        head = noc_head.detach().to(torch.float32).contiguous()
        b, c, hh, wh = head.shape
        if c != 3:
            raise ValueError('noc_head must be [B,3,Hh,Wh]')
        roi_hw = roi_hw.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty(b, 3, height, width, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            code = lib.posefit_resample_noc(_ptr(head), _ptr(roi_hw), b, hh, wh, height, width, _ptr(out), _stream(dev))
        _lib.check(code, 'posefit_resample_noc')
        ctx.save_for_backward(roi_hw)
        ctx.shape = (b, hh, wh, height, width)
        ctx.in_dtype = noc_head.dtype
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.lib()
        (roi_hw,) = ctx.saved_tensors
        b, hh, wh, height, width = ctx.shape
        g = grad_out.detach().to(torch.float32).contiguous()
        grad_head = torch.empty(b, 3, hh, wh, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            code = lib.posefit_resample_noc_backward(_ptr(g), _ptr(roi_hw), b, hh, wh, height, width, _ptr(grad_head),
                                                     _stream(g.device))
        _lib.check(code, 'posefit_resample_noc_backward')
        return grad_head.to(ctx.in_dtype), None, None, None


def resample_noc(noc_head, roi_hw, height: int, width: int):
    return ResampleNoc.apply(noc_head, roi_hw, height, width)


class BatchedPoses(NamedTuple):
    """Everything run_pose returns per instance (pose_estimation.py:401-412), for the whole batch, on the GPU."""
    global_rot: torch.Tensor     # [B,3,3] f64, scale embedded (:404-406)
    global_trans: torch.Tensor   # [B,3]
    global_scale: torch.Tensor   # [B]
    euler: torch.Tensor          # [B,3] XYZ Euler angles of the unscaled rotation (postprocess.py:158-160)
    world_box: torch.Tensor      # [B,8,3] sort_bbox-ordered world box of the depth points (:373-380)
    status: torch.Tensor         # [B] i32: 0 ok; 1 / 2 = the cases run_pose answers with 6 x None; 3 = NaN input
    scale: torch.Tensor          # [B]     camera-space fit, connected to autograd when pred_nocs requires grad:
    rot: torch.Tensor            # [B,3,3] the TRUE rotation (object -> camera is [s R | t], pose_estimation.py:401-403)
    trans: torch.Tensor          # [B,3]
    # per launch group -- ONE group (plain tensors) unless `bucket` is given, then lists with one entry per group:
    raw: object                  # PoseFitRaw of the fit (float64 records, inlier mask, winners)
    noc: object                  # [Bg,3,H,W] resampled NOC crops (differentiable w.r.t. the head output)
    crops: object                # Crops
    mask: object                 # [Bg,H,W] u8 correspondences that reached the fit (after the pre-filters)
    group_index: Optional[list] = None   # bucketed runs: instance indices (ascending) of every group


def _prepare(pred_nocs, depth_frames, inst_masks, boxes, frame_of, campose, kinv, gt_boxes, apply_statistical_filter,
             height, width):
    """gather -> resample -> GT clip -> the two outlier filters, for instances that share one H x W canvas."""
    from .function import clip_mask_to_box, statistical_outlier_mask
    crops = gather_crops(depth_frames, inst_masks, boxes, frame_of, height, width)
    noc = resample_noc(pred_nocs, crops.roi_hw, height, width)
    mask = crops.mask
    cam_index = frame_of if (campose is not None and torch.as_tensor(campose).dim() == 3) else None
    if gt_boxes is not None and campose is not None:
        mask, _ = clip_mask_to_box(crops.depth, mask, crops.bbox_xy0, gt_boxes, campose, kinv, cam_index=cam_index)
    if apply_statistical_filter:
        mask = statistical_outlier_mask(None, crops.depth, mask, crops.bbox_xy0, kinv, source='depth')
        mask = statistical_outlier_mask(noc.detach(), crops.depth, mask, crops.bbox_xy0, kinv, source='noc')
    return crops, noc, mask, cam_index


def _draw_sample_idx(counts, n_iterations: int, n_samples: int) -> torch.Tensor:
    """The reference's draws (pose_utils.py:73), instance after instance, nothing for an empty instance."""
    idx = np.zeros((len(counts), n_iterations, n_samples), dtype=np.int32)
    for i, n in enumerate(counts):
        if n > 0:
            idx[i] = np.random.randint(int(n), size=(n_iterations, n_samples))
    return torch.from_numpy(idx)


def _fit_on_reference_stream(counts, n_iterations: int, n_samples: int, run):
    """Fit with np.random draws and leave the global stream where the reference's per-instance loop leaves it.

    The reference draws inside getRANSACInliers' loop and stops drawing at the early-stop break (pose_utils.py:73,
    :80-81), so the stream position of instance i+1 depends on how many iterations instance i ran.  Draw for every
    instance assuming no early stop (true for all but near-perfect objects), fit, read back the iterations each fit says
    the reference would have run; if an instance stopped early, replay the stream from the state before the first draw
    -- instances up to it consume exactly their iterations, the later ones are drawn afresh -- and fit again.
    run(sample_idx [B,n_it,n_s] CPU int32) -> (result, iterations [B] numpy int)."""
    counts = np.asarray(counts)
    state0 = np.random.get_state()
    idx = _draw_sample_idx(counts, n_iterations, n_samples)
    assumed = np.where(counts > 0, n_iterations, 0)
    while True:
        out, iters = run(idx)
        late = np.nonzero((counts > 0) & (np.asarray(iters) != assumed))[0]
        if late.size == 0:
            return out
        k = int(late[0])
        assumed[k] = int(iters[k])
        assumed[k + 1:] = np.where(counts[k + 1:] > 0, n_iterations, 0)
        np.random.set_state(state0)
        rows = idx.numpy()
        for i, n in enumerate(counts):
            if n > 0 and assumed[i] > 0:
                rows[i, :assumed[i]] = np.random.randint(int(n), size=(int(assumed[i]), n_samples))


def _fit(noc, crops, mask, kinv, campose, cam_index, sample_idx, ransac, bits=False):
    from .function import PoseFitFull, PoseFitRaw, pose_epilogue, REF_COMPAT, SAMPLES_ARE_BITS
    # one forward feeds both autograd (the reference detaches here, postprocess.py:151; we do not have to) and the epilogue
    scale, rot, trans, inl, status, n_valid, pose64, winner, ctx64 = PoseFitFull.apply(
        noc, crops.depth, mask, crops.bbox_xy0, kinv, sample_idx if ransac else None, 1.0,
        REF_COMPAT | (SAMPLES_ARE_BITS if bits else 0))
    raw = PoseFitRaw(pose64, ctx64, status, n_valid, inl if ransac else None, winner if ransac else None)
    epi = pose_epilogue(raw, crops.depth, mask, crops.bbox_xy0, kinv, campose=campose, cam_index=cam_index)
    return BatchedPoses(epi.global_rot, epi.global_trans, epi.global_scale, epi.euler, epi.world_box, status,
                        scale, rot, trans, raw, noc, crops, mask)


def _canvas(boxes: torch.Tensor):
    bh = int((boxes[:, 3] - boxes[:, 1]).max()) if boxes.numel() else 1
    bw = int((boxes[:, 2] - boxes[:, 0]).max()) if boxes.numel() else 1
    return max(bh, 1), max((bw + 3) // 4 * 4, 4)


def run_pose_batched(pred_nocs, depth_frames, inst_masks, boxes_xyxy, frame_of: Optional[torch.Tensor] = None,
                     campose=None, kinv=None, gt_boxes=None, ransac: bool = True, n_iterations: int = 100,
                     n_samples: int = 10, apply_statistical_filter: bool = True, sample_idx=None,
                     height: Optional[int] = None, width: Optional[int] = None,
                     bucket: Optional[int] = None, generator=None) -> BatchedPoses:
    """The per-instance loop of `postprocess_dets` (Detection/tracker/postprocess.py:131-165) -- roi_align of
    the NOC head output, depth / mask slicing, run_pose -- for ALL instances of a batch of frames at once.

    pred_nocs [B,3,Hh,Wh] head outputs; depth_frames [F,FH,FW]; inst_masks [B,FH,FW]; boxes_xyxy [B,4];
    frame_of [B] frame index of each instance (None: one frame); campose [4,4] / [F,4,4] camera-to-world
    (None keeps camera space, as run_pose_office); gt_boxes [B,8,3] enables the clean_depth clip (:293-299).
    RANSAC indices are drawn from the global `np.random` stream exactly as the per-instance drop-in does
    (instance after instance, `randint(N_i, size=(n_iterations, n_samples))`, nothing for an empty instance)
    unless `sample_idx` [B,n_hyp,n_samp] is given; the only host round trip is the B correspondence counts
    that `randint` needs.  sample_idx='device' removes that round trip too: the draws are made on the GPU (torch's
    Philox generator, `generator=`) as uniform 32-bit values that the kernels map to floor(u N / 2^32) once they know N
    -- an opt-in, NOT numpy's stream.  Early stop and the global stream: the reference draws 10 indices per iteration
    and stops drawing at its early-stop break (pose_utils.py:73-81).  The indices here are drawn up front, but the fit
    reports how many iterations the reference would have run; after an early stop -- total residual below PassT / 100,
    i.e. a near-perfect object -- the stream is replayed so that every later instance draws from the position the
    reference would draw from, and the call leaves np.random where the reference's loop leaves it
    (`_fit_on_reference_stream`; one extra fit per early-stopping instance, none otherwise).
    Instances are padded to one H x W (default: the largest box, width rounded up to 4).

    bucket=k: boxes of very different sizes make that padding the dominant traffic (a 24x24 box on a 160x200 canvas
    reads 55x its own bytes).  With `bucket` the instances are grouped by box size rounded up to multiples of k and
    every group runs on its own canvas (one launch sequence per group, still one host round trip and the same draw
    order); the per-instance outputs come back in instance order, the per-group tensors (`raw`, `noc`, `crops`,
    `mask`) as lists with `group_index` naming the instances of each group."""
    boxes = torch.as_tensor(boxes_xyxy)
    if bucket and int(boxes.shape[0]) > 0:
        return _run_pose_bucketed(pred_nocs, depth_frames, inst_masks, boxes, frame_of, campose, kinv, gt_boxes, ransac,
                                  n_iterations, n_samples, apply_statistical_filter, sample_idx, int(bucket), generator)
    if height is None or width is None:
        ch, cw = _canvas(boxes)
        height, width = height or ch, width or cw
    elif not boxes.is_cuda and boxes.numel():
        # explicit canvas + boxes on the host: check here (device-resident boxes are checked by the gather kernel, which
        # emits an instance whose box exceeds the canvas as EMPTY -- status 1 -- rather than cropping it)
        bh, bw = int((boxes[:, 3] - boxes[:, 1]).max()), int((boxes[:, 2] - boxes[:, 0]).max())
        if bh > height or bw > width:
            raise ValueError(f'canvas {height}x{width} is smaller than the largest box ({bh}x{bw})')
    crops, noc, mask, cam_index = _prepare(pred_nocs, depth_frames, inst_masks, boxes, frame_of, campose, kinv, gt_boxes,
                                           apply_statistical_filter, height, width)
    bits = isinstance(sample_idx, str)
    if bits:
        if sample_idx != 'device':
            raise ValueError("sample_idx must be a tensor, None or 'device'")
        from .function import device_sample_bits
        sample_idx = device_sample_bits(int(mask.shape[0]), n_iterations, n_samples, mask.device, generator)
    if ransac and sample_idx is None:
        counts = ((mask != 0) & (crops.depth > 0)).flatten(1).sum(1).cpu().numpy()     # host round trip: randint needs N

        def run(idx):
            out = _fit(noc, crops, mask, kinv, campose, cam_index, idx, ransac, False)
            return out, out.raw.ctx[:, 30].cpu().numpy().astype(np.int64)              # iterations the reference runs
        return _fit_on_reference_stream(counts, n_iterations, n_samples, run)
    return _fit(noc, crops, mask, kinv, campose, cam_index, sample_idx, ransac, bits)


def bucket_groups(boxes_xyxy, bucket: int) -> dict:
    """Instances grouped by canvas: {(H, W): [instance indices, ascending]} with H = box height rounded up to a
    multiple of `bucket`, W = box width rounded up to a multiple of `bucket` and then of 4 (the vector loaders)."""
    boxes = torch.as_tensor(boxes_xyxy).cpu().to(torch.int64)
    hw = torch.stack([boxes[:, 3] - boxes[:, 1], boxes[:, 2] - boxes[:, 0]], dim=1).clamp_(min=1)
    key = (hw + bucket - 1) // bucket * bucket
    key[:, 1] = (key[:, 1] + 3) // 4 * 4
    groups = {}
    for i in range(int(boxes.shape[0])):
        groups.setdefault((int(key[i, 0]), int(key[i, 1])), []).append(i)
    return groups


def _run_pose_bucketed(pred_nocs, depth_frames, inst_masks, boxes, frame_of, campose, kinv, gt_boxes, ransac,
                       n_iterations, n_samples, apply_statistical_filter, sample_idx, bucket: int,
                       generator=None) -> BatchedPoses:
    b = int(boxes.shape[0])
    bits = isinstance(sample_idx, str)
    if bits:
        if sample_idx != 'device':
            raise ValueError("sample_idx must be a tensor, None or 'device'")
        from .function import device_sample_bits
        sample_idx = device_sample_bits(b, n_iterations, n_samples, pred_nocs.device, generator)
    groups = bucket_groups(boxes, bucket)
    dev = pred_nocs.device

    def take(t, idx):
        return None if t is None else torch.as_tensor(t).to(dev)[idx]

    per_object_kinv = kinv is not None and torch.as_tensor(kinv).dim() == 3
    prepared = []
    for (gh, gw), members in sorted(groups.items()):
        idx = torch.tensor(members, dtype=torch.long, device=dev)
        k_g = take(kinv, idx) if per_object_kinv else kinv
        crops, noc, mask, cam_index = _prepare(pred_nocs[idx], depth_frames, take(inst_masks, idx), take(boxes, idx),
                                               take(frame_of, idx), campose, k_g, take(gt_boxes, idx),
                                               apply_statistical_filter, gh, gw)
        prepared.append((idx, k_g, crops, noc, mask, cam_index))
    def fit_groups(sidx):
        sidx = None if sidx is None else torch.as_tensor(sidx).to(dev)
        return [_fit(noc, crops, mask, k_g, campose, cam_index, sidx[idx] if ransac else None, ransac, bits)
                for idx, k_g, crops, noc, mask, cam_index in prepared]

    if ransac and sample_idx is None:
        counts = torch.zeros(b, dtype=torch.long, device=dev)
        for idx, _, crops, _, mask, _ in prepared:
            counts[idx] = ((mask != 0) & (crops.depth > 0)).flatten(1).sum(1)

        def run(sidx):
            parts = fit_groups(sidx)
            iters = torch.zeros(b, dtype=torch.float64, device=dev)
            for pr, q in zip(prepared, parts):
                iters[pr[0]] = q.raw.ctx[:, 30]
            return parts, iters.cpu().numpy().astype(np.int64)
        parts = _fit_on_reference_stream(counts.cpu().numpy(), n_iterations, n_samples, run)   # host round trip: randint needs N
    else:
        parts = fit_groups(sample_idx)
    order = torch.cat([pr[0] for pr in prepared]) if prepared else torch.zeros(0, dtype=torch.long, device=dev)
    inv = torch.argsort(order)

    def merged(name):
        return torch.cat([getattr(q, name) for q in parts])[inv]
    fields = ('global_rot', 'global_trans', 'global_scale', 'euler', 'world_box', 'status', 'scale', 'rot', 'trans')
    return BatchedPoses(*[merged(f) for f in fields], [q.raw for q in parts], [q.noc for q in parts],
                        [q.crops for q in parts], [q.mask for q in parts], [pr[0] for pr in prepared])
