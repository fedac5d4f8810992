"""Gradient oracle -- TEST INFRASTRUCTURE.  The reference never back-propagates through the pose
fit (it detaches the NOC patch first, Detection/tracker/postprocess.py:151, and returns via
torch.from_numpy, :162-165), so there is no reference gradient to copy.  This is a float64 torch
transcription of estimateSimilarityUmeyama (PoseEst/pose_utils.py:16-61) whose gradients come
from torch autograd.  It is PINNED two ways: its forward is checked against the NumPy oracle on
every test input, and its gradients against central finite differences of the REAL reference
functions (tests/golden/grad_fd.npz, written by oracle/gen_golden_grad.py;
tests/test_math_host.py::test_gradients_match_reference_finite_differences).
"""
from __future__ import annotations

import torch


def umeyama_torch(src: torch.Tensor, dst: torch.Tensor, weights: torch.Tensor | None = None):
    """src, dst: [N,3] float64 (requires_grad allowed); weights: [N] 0/1 selection (the RANSAC
    inlier mask, treated as a constant).  Returns (s, R, t) with R the TRUE rotation (the
    reference reports R^T, pose_utils.py:44)."""
    if weights is None:
        weights = torch.ones(src.shape[0], dtype=src.dtype, device=src.device)
    w = weights / weights.sum()
    mu_s = (w[:, None] * src).sum(0)                                 # :23
    mu_d = (w[:, None] * dst).sum(0)                                 # :24
    cs, cd = src - mu_s, dst - mu_d                                  # :27-28
    cov = (w[:, None] * cd).T @ cs                                   # :30
    u, d, vh = torch.linalg.svd(cov)                                 # :38
    sign = torch.sign(torch.linalg.det(u) * torch.linalg.det(vh))    # :39-42
    fix = torch.stack([torch.ones_like(sign), torch.ones_like(sign), sign])
    rot = u @ torch.diag(fix) @ vh                                   # true R ; Rotation = rot.T (:44)
    var_s = (w[:, None] * cs * cs).sum()                             # :46
    s = (d * fix).sum() / var_s                                      # :47-48
    t = mu_d - s * (rot @ mu_s)                                      # :55
    return s, rot, t


def fit_gradients(src, dst, weights, g_s, g_R, g_t):
    """d<loss>/d(src, dst) for loss = g_s*s + <g_R, R> + <g_t, t>."""
    src = src.clone().requires_grad_(True)
    dst = dst.clone().requires_grad_(True)
    s, rot, t = umeyama_torch(src, dst, weights)
    loss = g_s * s + (g_R * rot).sum() + (g_t * t).sum()
    loss.backward()
    return src.grad, dst.grad, (s.detach(), rot.detach(), t.detach())
