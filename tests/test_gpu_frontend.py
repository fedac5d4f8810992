"""Batched front end (gather_crops, resample_noc) against torchvision.ops.roi_align -- the function
detectron2.layers.roi_align wraps and the reference calls per instance (postprocess.py:141-147) --
and NumPy slicing (pose_estimation.py:260-262, :290)."""
import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import posefit_oracle as po

pytestmark = pytest.mark.gpu
tv_ops = pytest.importorskip('torchvision.ops')


@pytest.fixture(scope='module')
def pf():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return load_pkg()


def _reference_patch(head_i, h, w):
    """postprocess.py:141-147 for one instance (CPU)."""
    box = [torch.tensor([[0.0, 0.0, float(head_i.shape[1]), float(head_i.shape[2])]])]
    return tv_ops.roi_align(head_i[None], box, output_size=(h, w), aligned=True)[0]


def test_resample_noc_matches_roi_align(pf):
    gen = torch.Generator().manual_seed(3)
    sizes = [(64, 64), (20, 33), (112, 90), (3, 5), (28, 28), (57, 14), (1, 1), (100, 7)]
    b, H, W = len(sizes), 112, 96
    head = torch.rand(b, 3, 28, 28, generator=gen)
    roi_hw = torch.tensor(sizes, dtype=torch.int32)
    got = pf.resample_noc(head.cuda(), roi_hw.cuda(), H, W).cpu()
    for i, (h, w) in enumerate(sizes):
        want = _reference_patch(head[i], h, w)
        np.testing.assert_allclose(got[i, :, :h, :w].numpy(), want.numpy(), rtol=0, atol=2e-7)
        assert float(got[i, :, h:, :].abs().max() if h < H else 0.0) == 0.0
        assert float(got[i, :, :, w:].abs().max() if w < W else 0.0) == 0.0


def test_resample_noc_backward_matches_autograd(pf):
    gen = torch.Generator().manual_seed(4)
    sizes = [(64, 64), (20, 33), (9, 40), (80, 96)]
    b, H, W = len(sizes), 80, 96
    head = torch.rand(b, 3, 28, 28, generator=gen)
    g = torch.randn(b, 3, H, W, generator=gen)
    hc = head.cuda().requires_grad_(True)
    out = pf.resample_noc(hc, torch.tensor(sizes, dtype=torch.int32).cuda(), H, W)
    out.backward(g.cuda())
    for i, (h, w) in enumerate(sizes):
        hi = head[i].clone().requires_grad_(True)
        _reference_patch(hi, h, w).backward(g[i, :, :h, :w])
        np.testing.assert_allclose(hc.grad[i].cpu().numpy(), hi.grad.numpy(), rtol=1e-4, atol=1e-5)


def test_gather_crops_matches_slicing(pf):
    rng = np.random.default_rng(5)
    F, FH, FW, H, W = 3, 240, 320, 72, 80
    depth = rng.uniform(0.5, 6.0, size=(F, FH, FW)).astype(np.float32)
    boxes = np.array([[10, 20, 74, 84], [100, 50, 180, 120], [300, 200, 320, 240], [0, 0, 5, 3], [50, 60, 50, 90],
                      [200, 100, 290, 190]], dtype=np.int32)          # the last one is larger than HxW
    b = boxes.shape[0]
    frame_of = np.array([0, 1, 2, 0, 1, 2], dtype=np.int32)
    masks = rng.uniform(size=(b, FH, FW)) < 0.6
    c = pf.gather_crops(torch.from_numpy(depth).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(boxes).cuda(),
                        torch.from_numpy(frame_of).cuda(), H, W)
    for i in range(b):
        x0, y0, x1, y1 = boxes[i]
        h, w = y1 - y0, x1 - x0
        if h > H or w > W:
            h = w = 0            # a box beyond the canvas is emitted EMPTY (never cropped: NOC would be resized to the wrong size)
        assert c.roi_hw[i].cpu().tolist() == [h, w] and c.bbox_xy0[i].cpu().tolist() == [x0, y0]
        want_d = np.zeros((H, W), dtype=np.float32)
        want_m = np.zeros((H, W), dtype=np.uint8)
        want_d[:h, :w] = depth[frame_of[i], y0:y0 + h, x0:x0 + w]
        want_m[:h, :w] = masks[i, y0:y0 + h, x0:x0 + w]
        np.testing.assert_array_equal(c.depth[i].cpu().numpy(), want_d)
        np.testing.assert_array_equal(c.mask[i].cpu().numpy(), want_m)
    # with the boxes on the host, run_pose_batched refuses an explicit canvas that is too small
    with pytest.raises(ValueError):
        pf.run_pose_batched(torch.rand(b, 3, 28, 28, device='cuda'), torch.from_numpy(depth).cuda(),
                            torch.from_numpy(masks).cuda(), torch.from_numpy(boxes), torch.from_numpy(frame_of).cuda(),
                            height=H, width=W, ransac=False, apply_statistical_filter=False)


def test_batched_pipeline_equals_per_instance_reference_flow(pf):
    """head output + frame depth + instance masks + boxes -> poses, in three launches, equals the
    per-instance flow of postprocess.py:131-152 + run_pose (oracle, no filters, plain fit)."""
    rng = np.random.default_rng(6)
    gen = torch.Generator().manual_seed(6)
    b, FH, FW, H, W = 5, 240, 320, 64, 72
    boxes = np.array([[40, 30, 100, 94], [150, 60, 222, 110], [10, 150, 60, 214], [200, 120, 264, 180], [90, 90, 131, 123]],
                     dtype=np.int32)
    head = torch.rand(b, 3, 28, 28, generator=gen)
    depth = rng.uniform(2.0, 5.0, size=(1, FH, FW)).astype(np.float32)
    masks = rng.uniform(size=(b, FH, FW)) < 0.7
    crops = pf.gather_crops(torch.from_numpy(depth).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(boxes).cuda(),
                            None, H, W)
    noc = pf.resample_noc(head.cuda(), crops.roi_hw, H, W)
    raw = pf.pose_fit_raw(noc, crops.depth, crops.mask, crops.bbox_xy0)
    for i in range(b):
        x0, y0, x1, y1 = boxes[i]
        patch = _reference_patch(head[i], y1 - y0, x1 - x0).permute(1, 2, 0).contiguous().numpy()   # HxWxC, :147
        noc_pts, depth_pts, _ = po.crop_correspondences(patch, depth[0], masks[i], boxes[i])
        o = po.pose_from_correspondences(noc_pts, depth_pts)
        pose = raw.pose[i].cpu().numpy()
        assert int(raw.status[i]) == o['status'] == 0 and int(raw.n_valid[i]) == o['n_valid']
        ang = np.degrees(np.linalg.norm(pose[1:10].reshape(3, 3) - o['R']) / np.sqrt(2))
        assert ang <= 1e-3
        np.testing.assert_allclose(pose[0], o['s'], rtol=1e-5)
        np.testing.assert_allclose(pose[10:13], o['t'], rtol=1e-5)


def test_run_pose_batched_equals_per_instance_drop_in(pf):
    """run_pose_batched (one call for all instances of two frames) vs the per-instance drop-in
    pose_estimation.run_pose fed with the reference-style roi_align patches, same np.random stream:
    global rotation / translation / scale and the world boxes agree; an empty instance is status 1."""
    rng = np.random.default_rng(8)
    gen = torch.Generator().manual_seed(8)
    FH, FW = 240, 320
    boxes = np.array([[40, 30, 100, 94], [150, 60, 222, 110], [10, 150, 60, 214], [200, 120, 264, 180], [90, 90, 131, 123],
                      [5, 5, 165, 125]], dtype=np.int32)                      # the last box: 120 x 160 (global-memory RANSAC mode)
    b = boxes.shape[0]
    frame_of = np.array([0, 0, 1, 1, 0, 1], dtype=np.int32)
    head = torch.rand(b, 3, 28, 28, generator=gen)
    # a consistent scene per instance: depth generated from the NOC patch through a known similarity
    depth = np.zeros((2, FH, FW), dtype=np.float32)
    masks = np.zeros((b, FH, FW), dtype=bool)
    k = po.motfront_intrinsics()
    for i in range(b):
        x0, y0, x1, y1 = boxes[i]
        h, w = y1 - y0, x1 - x0
        patch = _reference_patch(head[i], h, w).permute(1, 2, 0).numpy().astype(np.float64)
        rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        rot *= np.sign(np.linalg.det(rot))
        pts = 1.5 * (patch.reshape(-1, 3) - 0.5) @ rot.T + np.array([0.0, 0.0, -3.5])
        z = (-pts[:, 2]).reshape(h, w) + rng.normal(scale=0.01, size=(h, w))
        m = rng.uniform(size=(h, w)) < 0.8
        free = depth[frame_of[i], y0:y1, x0:x1] == 0
        depth[frame_of[i], y0:y1, x0:x1] = np.where(free, z, depth[frame_of[i], y0:y1, x0:x1]).astype(np.float32)
        masks[i, y0:y1, x0:x1] = m & free
    masks[4] = False                                                           # an empty instance -> 6 x None / status 1
    campose = np.tile(np.identity(4), (2, 1, 1))
    for f in range(2):
        campose[f, :3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        campose[f, :3, 3] = rng.normal(size=3)
    np.random.seed(1234)
    out = pf.run_pose_batched(head.cuda(), torch.from_numpy(depth).cuda(), torch.from_numpy(masks).cuda(),
                              torch.from_numpy(boxes).cuda(), torch.from_numpy(frame_of).cuda(),
                              campose=torch.from_numpy(campose))
    np.random.seed(1234)
    for i in range(b):
        x0, y0, x1, y1 = boxes[i]
        patch = _reference_patch(head[i], y1 - y0, x1 - x0).permute(1, 2, 0).contiguous()
        ref = pf.pose_estimation.run_pose(patch.cuda(), depth[frame_of[i]], campose[frame_of[i]],
                                          torch.from_numpy(masks[i]).cuda(), tuple(int(v) for v in boxes[i]))
        if ref[0] is None:
            assert int(out.status[i]) in (1, 2)
            continue
        assert int(out.status[i]) == 0
        np.testing.assert_allclose(out.global_rot[i].cpu().numpy(), ref[0], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out.global_trans[i].cpu().numpy(), ref[1], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(float(out.global_scale[i]), ref[2], rtol=1e-5)
        np.testing.assert_allclose(out.world_box[i].cpu().numpy(), ref[3], rtol=1e-6, atol=1e-6)
    assert int(out.status[4]) == 1


def test_run_pose_batched_follows_the_reference_stream_after_an_early_stop(pf):
    """An instance whose first hypothesis already scores below StopT ends the reference's RANSAC loop -- and its
    np.random draws -- after ONE iteration (pose_utils.py:73, :80-81), so every later instance draws from an earlier
    stream position than an up-front draw of 100 x 10 per instance would give it.  run_pose_batched and the per-instance
    drop-in both follow that: same poses, same iteration counts, and np.random ends where the lazy draws end."""
    rng = np.random.default_rng(18)
    gen = torch.Generator().manual_seed(18)
    FH, FW = 240, 320
    boxes = np.array([[40, 30, 100, 94], [150, 60, 178, 88], [10, 150, 60, 214], [200, 120, 264, 180]], dtype=np.int32)
    b = boxes.shape[0]
    frame_of = np.zeros(b, dtype=np.int32)
    head = torch.rand(b, 3, 28, 28, generator=gen)
    depth = np.zeros((1, FH, FW), dtype=np.float32)
    masks = np.zeros((b, FH, FW), dtype=bool)
    kinv = np.linalg.inv(po.motfront_intrinsics())
    for i in range(b):
        x0, y0, x1, y1 = boxes[i]
        h, w = y1 - y0, x1 - x0
        if i == 1:
            # 28 x 28 box: roi_align of the 28 x 28 head is the identity, so the head IS the NOC patch.  Exact similarity
            # with identity rotation (only those score ~0): noc = (p - t) / s + 0.5 for the back-projected p
            vv, uu = np.meshgrid(np.arange(y0, y1), np.arange(x0, x1), indexing='ij')
            z = (3.0 + 0.2 * np.sin(uu / 5.0) + 0.1 * np.cos(vv / 7.0)).astype(np.float32)
            rays = np.stack([uu, vv, np.ones_like(uu)], axis=-1).astype(np.float64) @ kinv.T
            pts = rays * z[..., None].astype(np.float64) * np.array([1.0, -1.0, -1.0])       # pose_estimation.py:34-41
            noc = (pts - np.array([0.1, -0.2, -3.0])) / 2.0 + 0.5
            head[i] = torch.from_numpy(noc).permute(2, 0, 1).to(torch.float32)
            m = np.ones((h, w), dtype=bool)
        else:
            patch = _reference_patch(head[i], h, w).permute(1, 2, 0).numpy().astype(np.float64)
            rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            rot *= np.sign(np.linalg.det(rot))
            pts = 1.5 * (patch.reshape(-1, 3) - 0.5) @ rot.T + np.array([0.0, 0.0, -3.5])
            z = ((-pts[:, 2]).reshape(h, w) + rng.normal(scale=0.01, size=(h, w))).astype(np.float32)
            m = rng.uniform(size=(h, w)) < 0.8
        depth[0, y0:y1, x0:x1] = z
        masks[i, y0:y1, x0:x1] = m
    campose = np.identity(4)
    campose[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    campose[:3, 3] = rng.normal(size=3)

    np.random.seed(4321)
    out = pf.run_pose_batched(head.cuda(), torch.from_numpy(depth).cuda(), torch.from_numpy(masks).cuda(),
                              torch.from_numpy(boxes).cuda(), torch.from_numpy(frame_of).cuda(),
                              campose=torch.from_numpy(campose), apply_statistical_filter=False)
    next_batched = np.random.randint(2 ** 31 - 1, size=4)
    iters = pf.ransac_iterations(out.raw).cpu().numpy()
    counts = out.raw.n_valid.cpu().numpy()
    assert iters[1] == 1 and (np.delete(iters, 1) == 100).all(), iters

    # the lazy draws of the reference, replayed: iteration count x 10 draws per instance, in instance order
    np.random.seed(4321)
    for i in range(b):
        np.random.randint(int(counts[i]), size=(int(iters[i]), 10))
    assert np.array_equal(np.random.randint(2 ** 31 - 1, size=4), next_batched)

    # the per-instance drop-in, same seed: same stream, same poses
    monkey = pf.pose_estimation.APPLY_STATISTICAL_FILTER
    pf.pose_estimation.APPLY_STATISTICAL_FILTER = False
    try:
        np.random.seed(4321)
        for i in range(b):
            x0, y0, x1, y1 = boxes[i]
            patch = _reference_patch(head[i], y1 - y0, x1 - x0).permute(1, 2, 0).contiguous()
            ref = pf.pose_estimation.run_pose(patch.cuda(), depth[0], campose, torch.from_numpy(masks[i]).cuda(),
                                              tuple(int(v) for v in boxes[i]))
            assert ref[0] is not None and int(out.status[i]) == 0, i
            np.testing.assert_allclose(out.global_rot[i].cpu().numpy(), ref[0], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(out.global_trans[i].cpu().numpy(), ref[1], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(float(out.global_scale[i]), ref[2], rtol=1e-5)
        assert np.array_equal(np.random.randint(2 ** 31 - 1, size=4), next_batched)
    finally:
        pf.pose_estimation.APPLY_STATISTICAL_FILTER = monkey
    # the early stopper recovered its similarity: scale 2, identity rotation
    np.testing.assert_allclose(float(out.scale[1]), 2.0, rtol=1e-5)
    np.testing.assert_allclose(out.rot[1].cpu().numpy(), np.identity(3), atol=1e-5)


def test_run_pose_batched_is_differentiable_to_the_head_output(pf):
    """Gradients of a pose loss reach the NOC head output through the fit and the resample -- the end-to-end
    path the reference cuts at postprocess.py:151 -- and equal the explicit composition resample -> pose_fit."""
    gen = torch.Generator().manual_seed(9)
    rng = np.random.default_rng(9)
    boxes = torch.tensor([[40, 30, 100, 94], [150, 60, 222, 110], [10, 150, 60, 214]], dtype=torch.int32)
    b = boxes.shape[0]
    head = torch.rand(b, 3, 28, 28, generator=gen).cuda().requires_grad_(True)
    depth = torch.from_numpy(rng.uniform(2.0, 5.0, size=(1, 240, 320)).astype(np.float32)).cuda()
    masks = torch.from_numpy(rng.uniform(size=(b, 240, 320)) < 0.7).cuda()
    out = pf.run_pose_batched(head, depth, masks, boxes.cuda(), ransac=False, apply_statistical_filter=False)
    assert out.scale.requires_grad and int((out.status != 0).sum()) == 0
    w_s, w_r, w_t = torch.randn(b, device='cuda'), torch.randn(b, 3, 3, device='cuda'), torch.randn(b, 3, device='cuda')
    ((out.scale * w_s).sum() + (out.rot * w_r).sum() + (out.trans * w_t).sum()).backward()
    g1 = head.grad.clone()
    head2 = head.detach().clone().requires_grad_(True)
    noc = pf.resample_noc(head2, out.crops.roi_hw, out.noc.shape[2], out.noc.shape[3])
    s2, r2, t2, _, _, _ = pf.pose_fit(noc, out.crops.depth, out.crops.mask, out.crops.bbox_xy0)
    ((s2 * w_s).sum() + (r2 * w_r).sum() + (t2 * w_t).sum()).backward()
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    # (the resample adjoint accumulates with float atomics: equal up to summation order)
    assert float((g1 - head2.grad).abs().max()) <= 1e-5 * float(g1.abs().max())


def test_run_pose_batched_bucketed_equals_single_canvas(pf):
    """bucket=k groups the instances by box size (one canvas per group instead of padding everything to the largest
    box): same draws, same statuses / winners / inlier decisions, poses equal to rounding, gradients to the head
    output equal, per-instance outputs in instance order."""
    rng = np.random.default_rng(21)
    gen = torch.Generator().manual_seed(21)
    FH, FW = 240, 320
    boxes = np.array([[40, 30, 64, 50], [150, 60, 222, 110], [10, 150, 60, 214], [200, 120, 264, 180], [90, 90, 131, 123],
                      [5, 5, 165, 125], [100, 10, 124, 33], [230, 20, 300, 100]], dtype=np.int32)
    b = boxes.shape[0]
    frame_of = np.array([0, 0, 1, 1, 0, 1, 0, 1], dtype=np.int32)
    head = torch.rand(b, 3, 28, 28, generator=gen)
    depth = np.zeros((2, FH, FW), dtype=np.float32)
    masks = np.zeros((b, FH, FW), dtype=bool)
    for i in range(b):
        x0, y0, x1, y1 = boxes[i]
        h, w = y1 - y0, x1 - x0
        patch = _reference_patch(head[i], h, w).permute(1, 2, 0).numpy().astype(np.float64)
        rot = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        rot *= np.sign(np.linalg.det(rot))
        pts = 1.5 * (patch.reshape(-1, 3) - 0.5) @ rot.T + np.array([0.0, 0.0, -3.5])
        z = (-pts[:, 2]).reshape(h, w) + rng.normal(scale=0.01, size=(h, w))
        free = depth[frame_of[i], y0:y1, x0:x1] == 0
        depth[frame_of[i], y0:y1, x0:x1] = np.where(free, z, depth[frame_of[i], y0:y1, x0:x1]).astype(np.float32)
        masks[i, y0:y1, x0:x1] = (rng.uniform(size=(h, w)) < 0.8) & free
    masks[6] = False                                                           # an empty instance
    campose = np.tile(np.identity(4), (2, 1, 1))
    for f in range(2):
        campose[f, :3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        campose[f, :3, 3] = rng.normal(size=3)
    args = (torch.from_numpy(depth).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(boxes).cuda(),
            torch.from_numpy(frame_of).cuda())
    outs = []
    for bucket in (None, 32):
        np.random.seed(77)
        outs.append(pf.run_pose_batched(head.cuda(), *args, campose=torch.from_numpy(campose), bucket=bucket))
    one, many = outs
    assert len(many.group_index) > 2 and sorted(torch.cat(many.group_index).tolist()) == list(range(b))
    assert one.status.tolist() == many.status.tolist() and int(one.status[6]) == 1
    ok = (one.status == 0).cpu().numpy()
    assert ok.sum() >= b - 2
    for name in ('global_rot', 'global_trans', 'global_scale', 'world_box', 'euler'):
        x, y = getattr(one, name).cpu().numpy()[ok], getattr(many, name).cpu().numpy()[ok]
        np.testing.assert_allclose(y, x, rtol=1e-9, atol=1e-9, err_msg=name)
    for g, idx in enumerate(many.group_index):                                 # winners and inlier masks, instance by instance
        for j, i in enumerate(idx.tolist()):
            assert int(many.raw[g].winner[j]) == int(one.raw.winner[i])
            h, w = boxes[i, 3] - boxes[i, 1], boxes[i, 2] - boxes[i, 0]
            assert torch.equal(many.raw[g].inlier_mask[j, :h, :w], one.raw.inlier_mask[i, :h, :w])
    # gradients to the head output (plain fit, no filters: the differentiable configuration)
    grads = []
    w_s, w_r, w_t = torch.randn(b, device='cuda'), torch.randn(b, 3, 3, device='cuda'), torch.randn(b, 3, device='cuda')
    for bucket in (None, 32):
        hd = head.cuda().requires_grad_(True)
        o = pf.run_pose_batched(hd, *args, ransac=False, apply_statistical_filter=False, bucket=bucket)
        ((o.scale * w_s).sum() + (o.rot * w_r).sum() + (o.trans * w_t).sum()).backward()
        grads.append(hd.grad.clone())
    assert torch.isfinite(grads[0]).all() and float(grads[0].abs().max()) > 0
    assert float((grads[0] - grads[1]).abs().max()) <= 1e-5 * float(grads[0].abs().max())


def test_head_fed_fit_equals_resample_then_fit(pf):
    """posefit_forward_head / posefit_backward_head (the roi_align resize fused into the fit's loaders)
    against the composition it replaces, resample_noc -> pose_fit: poses to rounding (the sampled NOC values are
    bit-identical, only the summation order differs), head and depth gradients to 1e-4 relative (float atomics).  Boxes
    smaller than the head map (several taps per pixel), larger, ragged canvas, empty instances, per-object intrinsics."""
    rng = np.random.default_rng(11)
    for (b, h, w, seed, per_obj) in [(48, 64, 64, 1, False), (24, 40, 52, 2, False), (12, 19, 27, 3, True), (6, 112, 112, 4, False)]:
        d = pf.synth.make_objects(b, h, w, seed=700 + seed, device='cuda', align_x0=1 if w % 4 else 4)
        head0 = torch.nn.functional.adaptive_avg_pool2d(d['noc'], 28).contiguous()
        roi = torch.from_numpy(np.stack([rng.integers(max(h // 4, 2), h + 1, size=b), rng.integers(max(w // 4, 2), w + 1, size=b)],
                                        axis=1).astype(np.int32)).cuda()
        roi[0] = torch.tensor([h, w], dtype=torch.int32)
        mask = d['mask'].clone()
        mask[1] = 0                                                       # an instance without correspondences
        kinv = None
        if per_obj:
            kinv = pf.default_kinv('cuda').repeat(b, 1, 1).contiguous()
            kinv[:, 0, 1] = 1e-4                                          # skewed: the general back-projection
        g = (torch.randn(b, device='cuda'), torch.randn(b, 3, 3, device='cuda'), torch.randn(b, 3, device='cuda'))

        def run(fused):
            head = head0.clone().requires_grad_(True)
            depth = d['depth'].clone().requires_grad_(True)
            if fused:
                s_, r_, t_, status, n_valid = pf.pose_fit_head(head, roi, depth, mask, d['bbox_xy0'], kinv)
            else:
                noc = pf.resample_noc(head, roi, h, w)
                s_, r_, t_, _, status, n_valid = pf.pose_fit(noc, depth, mask, d['bbox_xy0'], kinv)
            torch.autograd.backward((s_, r_, t_), g)
            return s_.detach(), r_.detach(), t_.detach(), status, n_valid, head.grad, depth.grad
        a, c = run(True), run(False)
        torch.cuda.synchronize()
        assert torch.equal(a[3], c[3]) and torch.equal(a[4], c[4])
        assert int(a[3][1]) == 1
        for x, y in zip(a[:3], c[:3]):
            assert float((x - y).abs().max()) <= 2e-6 * max(float(y.abs().max()), 1.0)
        for x, y in zip(a[5:], c[5:]):
            assert float((x - y).abs().max()) <= 1e-4 * float(y.abs().max()), (h, w, float((x - y).abs().max()), float(y.abs().max()))


@pytest.mark.gpu
def test_bit_packed_masks_expand_to_the_byte_masks(pf):
    """The one-bit-per-pixel wire format (pf.pack_mask on the host, posefit_unpack_mask on the device) against
    numpy.unpackbits, on aligned and unaligned pixel counts and an output that is not 8-byte aligned."""
    rng = np.random.default_rng(77)
    for shape in [(5, 64, 64), (3, 7, 9), (1, 1, 1), (2, 33, 5), (40, 112, 112)]:
        m = (rng.random(shape) < 0.6)
        bits = pf.pack_mask(m)
        assert bits.numel() == (m.size + 7) // 8
        out = pf.unpack_mask(bits.cuda(), shape)
        assert out.dtype == torch.uint8 and tuple(out.shape) == shape
        assert np.array_equal(out.cpu().numpy(), m.astype(np.uint8))
    # an unaligned destination takes the byte path: expand into an offset view through the C ABI directly
    m = rng.random(1003) < 0.5
    bits = pf.pack_mask(m).cuda()
    buf = torch.full((1003 + 16,), 7, dtype=torch.uint8, device='cuda')
    lib = pf._lib.lib()
    code = lib.posefit_unpack_mask(bits.data_ptr(), 1003, buf.data_ptr() + 3, torch.cuda.current_stream().cuda_stream)
    assert code == 0
    torch.cuda.synchronize()
    got = buf.cpu().numpy()
    assert np.array_equal(got[3:1006], m.astype(np.uint8)) and (got[:3] == 7).all() and (got[1006:] == 7).all()
    # the fit on an unpacked mask is the fit on the byte mask
    d = pf.synth.make_objects(6, 64, 64, seed=91)
    t = {k: d[k].cuda() for k in ('noc', 'depth', 'mask', 'bbox_xy0')}
    a = pf.pose_fit_raw(t['noc'], t['depth'], t['mask'], t['bbox_xy0'])
    b = pf.pose_fit_raw(t['noc'], t['depth'], pf.unpack_mask(pf.pack_mask(d['mask']).cuda(), d['mask'].shape), t['bbox_xy0'])
    assert torch.equal(a.pose, b.pose) and torch.equal(a.status, b.status)
