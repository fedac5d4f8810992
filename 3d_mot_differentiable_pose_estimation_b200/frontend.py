"""Batched front end: what `postprocess_dets` does per instance before it calls run_pose
(Detection/tracker/postprocess.py:131-152), for a whole batch of instances in two launches.

  crops = gather_crops(depth_frames, inst_masks, boxes_xyxy, frame_of, H, W)
  noc   = resample_noc(noc_head, crops.roi_hw, H, W)          # differentiable w.r.t. noc_head
  out   = pose_fit(noc, crops.depth, crops.mask, crops.bbox_xy0, ...)
"""
from __future__ import annotations

from typing import NamedTuple, Optional

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
