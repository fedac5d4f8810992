"""Synthetic MOTFront-shaped pose-solver inputs.

One "object" is what the reference's `run_pose` sees for one detected instance
(PoseEst/pose_estimation.py:245-323) in the crop layout the CUDA path consumes:
a NOC crop `[3,h,w]` in [0,1], the depth crop `[h,w]` in metres, the instance mask crop
`[h,w]` and the crop's top-left corner inside the 240x320 frame.  The fixed MOTFront
camera (pose_estimation.py:269-288) relates pixels to rays.

Each object gets a random similarity (s, R, t); depth is a jittered cloud around the
object centre, NOC is the inverse transform of the back-projected depth point (+ noise,
clipped to [0,1]), so a consistent pose exists.  A fraction of pixels become gross
outliers (depth pushed metres away AFTER the NOC was derived), the mask is Bernoulli
with an empty border, and a few depths are zeroed (invalid).  Everything is built with
batched torch ops on `device`, in chunks, from a seeded `torch.Generator`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

FRAME_H, FRAME_W = 240, 320
FOCAL = 292.87803547399                      # pose_estimation.py:272-273


def motfront_intrinsics(height: int = FRAME_H, width: int = FRAME_W) -> torch.Tensor:
    """K of run_pose (pose_estimation.py:269-288), float64 [3,3] on the CPU."""
    return torch.tensor([[FOCAL, 0.0, width / 2 - 0.5],
                         [0.0, FOCAL, height / 2 - 0.5],
                         [0.0, 0.0, 1.0]], dtype=torch.float64)


def _random_rotations(n: int, gen: torch.Generator, device) -> torch.Tensor:
    q = torch.randn(n, 4, generator=gen, device=device, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).reshape(n, 3, 3)


def make_objects(n_objects: int, height: int = 64, width: int = 64, *, seed: int = 0,
                 device: str | torch.device = 'cpu', n_hyp: int = 0, n_samp: int = 10,
                 outlier_frac: float = 0.10, outlier_range=(8.0, 20.0), mask_fill: float = 0.7,
                 noc_noise: float = 0.01, zero_depth_frac: float = 0.02, border: int = 2,
                 align_x0: int = 4, chunk: int = 4096,
                 out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """Returns dict(noc[B,3,h,w] f32, depth[B,h,w] f32, mask[B,h,w] u8, bbox_xy0[B,2] i32,
    gt_scale[B] f64, gt_R[B,3,3] f64, gt_t[B,3] f64, n_valid[B] i32 and, when n_hyp>0,
    sample_idx[B,n_hyp,n_samp] i32 -- indices into the row-major list of valid pixels, the
    replacement for `np.random.randint(N, size=10)` of pose_utils.py:73)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    b, h, w = n_objects, height, width
    if out is None:
        out = dict(
            noc=torch.empty(b, 3, h, w, dtype=torch.float32, device=device),
            depth=torch.empty(b, h, w, dtype=torch.float32, device=device),
            mask=torch.empty(b, h, w, dtype=torch.uint8, device=device),
        )
    out['bbox_xy0'] = torch.empty(b, 2, dtype=torch.int32, device=device)
    out['gt_scale'] = torch.empty(b, dtype=torch.float64, device=device)
    out['gt_R'] = torch.empty(b, 3, 3, dtype=torch.float64, device=device)
    out['gt_t'] = torch.empty(b, 3, dtype=torch.float64, device=device)
    out['n_valid'] = torch.empty(b, dtype=torch.int32, device=device)
    if n_hyp > 0:
        out['sample_idx'] = torch.empty(b, n_hyp, n_samp, dtype=torch.int32, device=device)

    cx, cy = FRAME_W / 2 - 0.5, FRAME_H / 2 - 0.5
    jj = torch.arange(w, device=device, dtype=torch.float64)
    ii = torch.arange(h, device=device, dtype=torch.float64)
    f32 = dict(generator=gen, device=device, dtype=torch.float32)
    f64 = dict(generator=gen, device=device, dtype=torch.float64)
    for c0 in range(0, b, chunk):
        n = min(chunk, b - c0)
        sl = slice(c0, c0 + n)
        nx = max((FRAME_W - w) // align_x0, 0) + 1
        x0 = (torch.rand(n, **f64) * nx).floor().clamp_(max=nx - 1) * align_x0
        y0 = (torch.rand(n, **f64) * (FRAME_H - h + 1)).floor().clamp_(max=FRAME_H - h)
        d0 = 2.5 + 2.0 * torch.rand(n, **f64)
        rot = _random_rotations(n, gen, device)
        # rays of every crop pixel: (rx, ry) so that the point is (rx*z, -ry*z, -z)
        rx = ((x0[:, None] + jj[None, :]) - cx) / FOCAL             # [n, w]
        ry = ((y0[:, None] + ii[None, :]) - cy) / FOCAL             # [n, h]
        width_m = d0 * w / FOCAL
        diag_m = d0 * math.sqrt(h * h + w * w) / FOCAL
        scale = diag_m * (1.0 + 0.6 * torch.rand(n, **f64))
        z = d0[:, None, None] + (0.5 * width_m)[:, None, None] * (torch.rand(n, h, w, **f64) - 0.5)
        z = z.to(torch.float32)                                     # what the solver will read
        zd = z.to(torch.float64)
        rcx = ((x0 + (w - 1) / 2) - cx) / FOCAL
        rcy = ((y0 + (h - 1) / 2) - cy) / FOCAL
        t = torch.stack([rcx * d0, -rcy * d0, -d0], dim=1)          # [n,3] object centre
        p = torch.stack([rx[:, None, :] * zd, -ry[:, :, None] * zd, -zd], dim=1)   # [n,3,h,w]
        q = p - t[:, :, None, None]
        noc = torch.einsum('nji,njhw->nihw', rot, q) / scale[:, None, None, None] + 0.5   # R^T (y - t)/s
        noc = noc + noc_noise * torch.randn(n, 3, h, w, **f64)
        out['noc'][sl] = noc.clamp_(0.0, 1.0).to(torch.float32)
        del p, q, noc, zd
        is_out = torch.rand(n, h, w, **f32) < outlier_frac
        push = outlier_range[0] + (outlier_range[1] - outlier_range[0]) * torch.rand(n, h, w, **f32)
        z = torch.where(is_out, z + push, z)
        z = torch.where(torch.rand(n, h, w, **f32) < zero_depth_frac, torch.zeros_like(z), z)
        m = torch.rand(n, h, w, **f32) < mask_fill
        if border > 0:
            m[:, :border, :] = False
            m[:, h - border:, :] = False
            m[:, :, :border] = False
            m[:, :, w - border:] = False
        out['depth'][sl] = z
        out['mask'][sl] = m.to(torch.uint8)
        out['bbox_xy0'][sl, 0] = x0.to(torch.int32)
        out['bbox_xy0'][sl, 1] = y0.to(torch.int32)
        out['gt_scale'][sl] = scale
        out['gt_R'][sl] = rot
        out['gt_t'][sl] = t
        nv = (m & (z > 0)).flatten(1).sum(1)
        out['n_valid'][sl] = nv.to(torch.int32)
        if n_hyp > 0:
            u = torch.rand(n, n_hyp, n_samp, **f64)
            idx = (u * nv[:, None, None].to(torch.float64)).floor().to(torch.int64)
            idx = torch.minimum(idx, (nv[:, None, None] - 1).clamp_(min=0))
            out['sample_idx'][sl] = idx.to(torch.int32)
        del is_out, push, z, m
    return out
