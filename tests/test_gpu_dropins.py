"""The same-named drop-ins (pose_utils.py / pose_estimation.py of the package) against the golden
vectors written from the REAL reference functions of the same name."""
import contextlib
import os

import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import posefit_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pf():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return load_pkg()


def _golden_aabb_permutation():
    """Row permutation sort_bbox (the reference's own, pose_estimation.py:72-93) applies to the 8 corners of an
    axis-aligned box with positive extent on every axis, read off the golden vectors of the real function."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'sort_bbox.npz'))
    perms = []
    for k in range(int(g['n'])):
        if str(g[f'tag_{k}']) == 'aabb extent mask 7':
            src, dst = g[f'in_{k}'], g[f'out_{k}']
            perms.append([int(np.where((src == row).all(axis=1))[0][0]) for row in dst])
    assert len(perms) >= 4 and all(p == perms[0] for p in perms)
    return perms[0]


def _hom(p):
    return np.transpose(np.hstack([p, np.ones([p.shape[0], 1])]))


def rot_err_deg(ra, rb):
    return float(np.degrees(np.linalg.norm(ra - rb) / np.sqrt(2)))


@contextlib.contextmanager
def replay(idx):
    """np.random.randint(n, size=(n_iter, 10)) -> the stored indices (what the reference drew)."""
    real = np.random.randint

    def fake(high, size=None, **kw):
        # the up-front draw of every iteration, or -- after an early stop -- the drop-in's replay of the iterations
        # that ran (pose_utils.rewind_to_reference_stream): a prefix of the same rows
        assert tuple(size)[1:] == idx.shape[1:] and size[0] <= idx.shape[0] and idx.max() < high
        return idx[:size[0]].copy()

    np.random.randint = fake
    try:
        yield
    finally:
        np.random.randint = real


def test_estimateSimilarityUmeyama(pf, golden_dir):
    g = np.load(os.path.join(golden_dir, 'umeyama.npz'))
    for k, name in enumerate(str(n) for n in g['names']):
        if name in ('roundoff_variance', 'single_point'):
            continue
        scales, rot, trans, tf = pf.pose_utils.estimateSimilarityUmeyama(_hom(g[f'src_{k}']), _hom(g[f'dst_{k}']))
        assert rot_err_deg(rot, g[f'rotation_{k}']) < 1e-6, name
        np.testing.assert_allclose(scales, g[f'scales_{k}'], rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(trans, g[f'translation_{k}'], rtol=1e-9, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(tf, g[f'transform_{k}'], rtol=1e-8, atol=1e-8, err_msg=name)
    bad = np.ones((4, 6))
    bad[1, 2] = np.nan
    with pytest.raises(RuntimeError, match='NANs'):
        pf.pose_utils.estimateSimilarityUmeyama(bad, np.ones((4, 6)))


def test_evaluateModel(pf, golden_dir):
    g = np.load(os.path.join(golden_dir, 'evaluate.npz'))
    for k in range(int(g['count'])):
        res, ratio, idx = pf.pose_utils.evaluateModel(g[f'transform_{k}'], _hom(g[f'src_{k}']), _hom(g[f'dst_{k}']),
                                                      float(g[f'pass_{k}']))
        np.testing.assert_allclose(res, g[f'residual_{k}'], rtol=1e-12)
        assert ratio == float(g[f'ratio_{k}'])
        np.testing.assert_array_equal(idx, g[f'idx_{k}'])


def test_getRANSACInliers(pf, golden_dir):
    g = np.load(os.path.join(golden_dir, 'ransac_inliers.npz'))
    for k in range(int(g['count'])):
        src, dst, idx = g[f'src_{k}'], g[f'dst_{k}'], g[f'idx_{k}']
        pass_t = float(g[f'pass_{k}'])
        with replay(idx):
            s_in, d_in, ratio = pf.pose_utils.getRANSACInliers(_hom(src), _hom(dst), MaxIterations=idx.shape[0],
                                                               PassThreshold=pass_t, StopThreshold=pass_t / 100)
        np.testing.assert_array_equal(s_in[:3].T, g[f'src_in_{k}'])
        np.testing.assert_array_equal(d_in[:3].T, g[f'dst_in_{k}'])
        assert ratio == float(g[f'ratio_{k}'])


def test_estimateSimilarityTransform(pf, golden_dir, monkeypatch, capsys):
    g = np.load(os.path.join(golden_dir, 'ransac.npz'))
    for k in range(int(g['count'])):
        name, idx = str(g[f'name_{k}']), g[f'idx_{k}']
        monkeypatch.setattr(pf.pose_utils, 'N_ITERATIONS', idx.shape[0])
        with replay(idx):
            scales, rot, trans, tf = pf.pose_utils.estimateSimilarityTransform(g[f'src_{k}'], g[f'dst_{k}'])
        if not bool(g[f'ok_{k}']):
            assert scales is None and rot is None and trans is None and tf is None, name
            assert 'Small BestInlierRatio' in capsys.readouterr().out
            continue
        assert rot_err_deg(rot, g[f'rotation_{k}']) < 1e-6, name
        np.testing.assert_allclose(scales, g[f'scales_{k}'], rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(trans, g[f'translation_{k}'], rtol=1e-9, atol=1e-9, err_msg=name)


def test_drop_ins_leave_np_random_where_the_reference_leaves_it(pf, golden_dir):
    """tests/golden/rng_stream.npz: the real estimateSimilarityTransform on a seeded, unpatched global stream.  The
    reference draws inside its RANSAC loop and stops drawing at the early stop (pose_utils.py:73, :80-81); the drop-in
    draws up front and then rewinds by the iterations the kernel reports: same outputs, same iteration count, and the
    NEXT values of np.random are the reference's."""
    g = np.load(os.path.join(golden_dir, 'rng_stream.npz'))
    seen = set()
    for name in [str(n) for n in g['names']]:
        src, dst = g[name + '_src'], g[name + '_dst']
        seed, ra, iters = int(g[name + '_seed']), float(g[name + '_ratio_adapt']), int(g[name + '_iterations'])
        seen.add(iters)
        np.random.seed(seed)
        scales, rot, trans, tf = pf.pose_utils.estimateSimilarityTransform(src, dst, ratio_adapt=ra)
        nxt = np.random.randint(2 ** 31 - 1, size=4)
        assert np.array_equal(nxt, g[name + '_next']), (name, iters)
        assert rot_err_deg(rot, g[name + '_rotation']) < 1e-6, name
        np.testing.assert_allclose(scales, g[name + '_scales'], rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(trans, g[name + '_translation'], rtol=1e-9, atol=1e-9, err_msg=name)
        # the raw entry reports the iteration count itself
        np.random.seed(seed)
        idx = torch.from_numpy(np.random.randint(src.shape[0], size=(100, 10)).astype(np.int32))[None].cuda()
        planes = lambda a: torch.from_numpy(np.ascontiguousarray(a.T))[None].cuda()    # noqa: E731  [1,3,N]
        raw = pf.points_fit_raw(planes(src), planes(dst), sample_idx=idx, ratio_adapt=ra)
        assert int(pf.ransac_iterations(raw)[0]) == iters, name
    assert 100 in seen and 1 in seen and len(seen) >= 4          # no stop, stop at once, stops in the middle


@pytest.mark.parametrize('tag,h,w,b', [('small', 24, 32, 6), ('odd', 19, 27, 4)])
def test_backproject_transform_cam2world(pf, golden_dir, tag, h, w, b):
    g = np.load(os.path.join(golden_dir, 'frames.npz'))
    k_mat = g[f'{tag}_K']
    for i in range(b):
        x0, y0 = (int(v) for v in g[f'{tag}_bbox_xy0'][i])
        depth = np.zeros((240, 320))
        mask = np.zeros((240, 320), dtype=bool)
        depth[y0:y0 + h, x0:x0 + w] = g[f'{tag}_depth'][i]
        mask[y0:y0 + h, x0:x0 + w] = g[f'{tag}_mask'][i] != 0
        pts, (rows, cols) = pf.pose_estimation.backproject(depth, k_mat, mask)
        if int(g[f'{tag}_{i}_status']) == 1:
            assert pts.shape == (0, 3)
            continue
        np.testing.assert_array_equal(rows, g[f'{tag}_{i}_rows'])
        np.testing.assert_array_equal(cols, g[f'{tag}_{i}_cols'])
        np.testing.assert_allclose(pts, g[f'{tag}_{i}_pts'], rtol=1e-14, atol=1e-14)
        if int(g[f'{tag}_{i}_status']) == 0:
            cam = pf.pose_estimation.transform_pc(g[f'{tag}_{i}_ransac_scales'], g[f'{tag}_{i}_ransac_rotation'],
                                                  g[f'{tag}_{i}_ransac_translation'], g[f'{tag}_{i}_noc_pts'])
            np.testing.assert_allclose(cam, g[f'{tag}_{i}_transformed_pc'], rtol=1e-12, atol=1e-12)
            world = pf.pose_estimation.cam2world(cam, g[f'{tag}_{i}_campose'])
            np.testing.assert_allclose(world, g[f'{tag}_{i}_world_pc'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('h,w', [(48, 56), (150, 172)])
def test_run_pose_against_oracle(pf, monkeypatch, h, w):
    """run_pose end to end (minus the Open3D / GT-clip filters) vs the oracle's restatement of
    pose_estimation.py:256-290, :323, :359-367, :401-412 with the same replayed indices.  The 150x172
    box (25 800 px) is beyond the shared-memory staging: the RANSAC kernel's global-memory mode."""
    d = pf.synth.make_objects(3, h, w, seed=77, outlier_range=(25.0, 40.0))
    rng = np.random.default_rng(3)
    for i in range(3):
        x0, y0 = (int(v) for v in d['bbox_xy0'][i])
        depth = rng.uniform(1.0, 5.0, size=(240, 320)).astype(np.float32)      # clutter outside the box is ignored
        depth[y0:y0 + h, x0:x0 + w] = d['depth'][i].numpy()
        mask = rng.uniform(size=(240, 320)) < 0.5
        mask[y0:y0 + h, x0:x0 + w] = d['mask'][i].numpy() != 0
        noc_hwc = d['noc'][i].permute(1, 2, 0).contiguous()
        campose = np.identity(4)
        campose[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        campose[:3, 3] = rng.normal(size=3)
        bbox = (x0, y0, x0 + w, y0 + h)
        noc_pts, depth_pts, _ = po.crop_correspondences(noc_hwc.numpy(), depth, mask, bbox)
        keep = po.run_pose_filters(noc_pts, depth_pts)                  # the two statistical filters
        assert 100 < len(keep) < noc_pts.shape[0]
        noc_pts, depth_pts = noc_pts[keep], depth_pts[keep]
        idx = rng.integers(0, noc_pts.shape[0], size=(100, 10))
        ora = po.pose_from_correspondences(noc_pts, depth_pts, idx)
        with replay(idx):
            out = pf.pose_estimation.run_pose(noc_hwc.cuda(), depth, campose, torch.from_numpy(mask).cuda(), bbox)
        assert ora['status'] == 0
        g_rot, g_trans, g_scale, box, depth_world, world_pc = out
        want = campose @ po.object_to_camera(np.full(3, ora['s']), ora['rot_t'], ora['t'])
        np.testing.assert_allclose(g_rot, want[:3, :3], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(g_trans, want[:3, 3], rtol=1e-7, atol=1e-7)
        np.testing.assert_allclose(g_scale, ora['s'], rtol=1e-9)
        np.testing.assert_allclose(depth_world, po.camera_to_world(depth_pts, campose), rtol=1e-12, atol=1e-12)
        cam = po.apply_similarity(np.full(3, ora['s']), ora['rot_t'], ora['t'], noc_pts)
        np.testing.assert_allclose(world_pc, po.camera_to_world(cam, campose), rtol=1e-6, atol=1e-6)
        assert box.shape == (8, 3)
        np.testing.assert_allclose(box.min(0), depth_world.min(0))
        np.testing.assert_allclose(box.max(0), depth_world.max(0))
        # office variant: camera space, explicit intrinsics
        with replay(idx):
            out2 = pf.pose_estimation.run_pose_office(noc_hwc.cuda(), torch.from_numpy(depth)[None],
                                                      torch.from_numpy(po.motfront_intrinsics())[None],
                                                      torch.from_numpy(mask).cuda(), bbox)
        np.testing.assert_allclose(out2[1], ora['t'], rtol=1e-9)
        np.testing.assert_allclose(out2[0], ora['s'] * ora['R'], rtol=1e-7, atol=1e-7)
    # empty mask -> 6 x None (pose_estimation.py:361-362)
    none = pf.pose_estimation.run_pose(noc_hwc.cuda(), depth, campose, torch.zeros(240, 320, dtype=torch.bool).cuda(), bbox)
    assert all(v is None for v in none)


def _euler_xyz_blender(m):
    """mathutils.Matrix(m).to_euler() ('XYZ'), restated from Blender's mat3_normalized_to_eul2
    (mathutils 2.81.2 is not installed: unpinned)."""
    cy = np.hypot(m[0, 0], m[1, 0])
    if cy > 16 * 1.1920929e-07:
        e1 = np.array([np.arctan2(m[2, 1], m[2, 2]), np.arctan2(-m[2, 0], cy), np.arctan2(m[1, 0], m[0, 0])])
        e2 = np.array([np.arctan2(-m[2, 1], -m[2, 2]), np.arctan2(-m[2, 0], -cy), np.arctan2(-m[1, 0], -m[0, 0])])
    else:
        e1 = np.array([np.arctan2(-m[1, 2], m[1, 1]), np.arctan2(-m[2, 0], cy), 0.0])
        e2 = e1
    return e2 if np.abs(e1).sum() > np.abs(e2).sum() else e1


def test_batched_epilogue(pf):
    """posefit_epilogue == the per-object tail of run_pose (pose_estimation.py:367-412) + the Euler
    conversion of postprocess.py:158-160, for a batch with per-frame camera poses."""
    b, h, w, frames = 24, 40, 48, 3
    d = pf.synth.make_objects(b, h, w, seed=91)
    d['mask'][5] = 0                                       # one empty object
    rng = np.random.default_rng(9)
    campose = np.tile(np.identity(4), (frames, 1, 1))
    for f in range(frames):
        campose[f, :3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        campose[f, :3, 3] = rng.normal(size=3)
    cam_index = np.arange(b) % frames
    t = {k: d[k].cuda() for k in ('noc', 'depth', 'mask', 'bbox_xy0')}
    raw = pf.pose_fit_raw(t['noc'], t['depth'], t['mask'], t['bbox_xy0'])
    epi = pf.pose_epilogue(raw, t['depth'], t['mask'], t['bbox_xy0'], campose=torch.from_numpy(campose),
                           cam_index=torch.from_numpy(cam_index))
    ora = po.batch_pose(d['noc'].numpy(), d['depth'].numpy(), d['mask'].numpy(), d['bbox_xy0'].numpy())
    for i, o in enumerate(ora):
        if o['status'] != 0:
            assert float(epi.world_box[i].abs().max()) == 0.0
            continue
        cp = campose[cam_index[i]]
        want = cp @ po.object_to_camera(np.full(3, o['s']), o['rot_t'], o['t'])
        np.testing.assert_allclose(epi.global_rot[i].cpu().numpy(), want[:3, :3], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(epi.global_trans[i].cpu().numpy(), want[:3, 3], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(float(epi.global_scale[i]), o['s'], rtol=1e-10)
        unscaled = want[:3, :3] / np.linalg.norm(want[:3, :3], axis=0)
        np.testing.assert_allclose(epi.euler[i].cpu().numpy(), _euler_xyz_blender(unscaled), rtol=1e-8, atol=1e-8)
        x0, y0 = (int(v) for v in d['bbox_xy0'][i])
        fd = np.zeros((240, 320), dtype=np.float32)
        fm = np.zeros((240, 320), dtype=bool)
        fd[y0:y0 + h, x0:x0 + w] = d['depth'][i].numpy()
        fm[y0:y0 + h, x0:x0 + w] = d['mask'][i].numpy() != 0
        pts, _ = po.backproject_points(fd.astype(np.float64), po.motfront_intrinsics(), fm)
        world = po.camera_to_world(pts, cp)
        # corner order: the permutation the REAL sort_bbox applies to a generic axis-aligned box (tests/golden/sort_bbox.npz,
        # oracle/gen_golden_bbox.py); the order Open3D's get_box_points() hands it the corners in stays unpinned
        box = pf.pose_estimation._aabb_corners(world)[_golden_aabb_permutation()]
        np.testing.assert_allclose(epi.world_box[i].cpu().numpy(), box, rtol=1e-12, atol=1e-12)
        np.testing.assert_array_equal(box, pf.pose_estimation.sort_bbox(pf.pose_estimation._aabb_corners(world)))
    # camera-space variant (run_pose_office)
    epi2 = pf.pose_epilogue(raw, t['depth'], t['mask'], t['bbox_xy0'])
    i = 0
    np.testing.assert_allclose(epi2.global_rot[i].cpu().numpy(), ora[i]['s'] * ora[i]['R'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(epi2.global_trans[i].cpu().numpy(), ora[i]['t'], rtol=1e-10)


def test_gt_box_clip_mask(pf):
    """posefit_clip_mask == clean_depth (pose_estimation.py:107-134) with run_pose's '> 20' rule (:293-299)."""
    b, h, w = 10, 40, 48
    d = pf.synth.make_objects(b, h, w, seed=93)
    rng = np.random.default_rng(4)
    campose = np.tile(np.identity(4), (b, 1, 1))
    boxes = np.zeros((b, 8, 3))
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
    want_masks, want_kept = [], []
    for i in range(b):
        campose[i, :3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        campose[i, :3, 3] = rng.normal(size=3)
        x0, y0 = (int(v) for v in d['bbox_xy0'][i])
        fd = np.zeros((240, 320), dtype=np.float32)
        fm = np.zeros((240, 320), dtype=bool)
        fd[y0:y0 + h, x0:x0 + w] = d['depth'][i].numpy()
        fm[y0:y0 + h, x0:x0 + w] = d['mask'][i].numpy() != 0
        pts, (rows, cols) = po.backproject_points(fd.astype(np.float64), po.motfront_intrinsics(), fm)
        world = po.camera_to_world(pts, campose[i])
        half = world.std(axis=0) * (0.02 if i == 3 else rng.uniform(0.5, 1.5))     # object 3: almost nothing survives
        boxes[i] = (np.median(world, axis=0) + corners * half)[rng.permutation(8)]
        keep = po.clipped_correspondence_indices(pts, boxes[i], campose[i])
        m = np.zeros((h, w), dtype=np.uint8)
        m[rows[keep] - y0, cols[keep] - x0] = 1
        want_masks.append(m)
        want_kept.append(len(keep))
    got, kept = pf.clip_mask_to_box(d['depth'].cuda(), d['mask'].cuda(), d['bbox_xy0'].cuda(),
                                    torch.from_numpy(boxes), torch.from_numpy(campose))
    assert kept.cpu().tolist() == want_kept
    assert want_kept[3] == int(d['n_valid'][3])                    # fallback to the unclipped set
    assert any(k < int(n) for k, n in zip(want_kept, d['n_valid'].tolist()))
    for i in range(b):
        np.testing.assert_array_equal(got[i].cpu().numpy(), want_masks[i])


@pytest.mark.parametrize('h,w', [(48, 56), (13, 17), (46, 45), (64, 64), (90, 100)])
def test_statistical_outlier_mask(pf, h, w):
    """posefit_sor_mask vs the oracle restatement of Open3D's remove_statistical_outlier (unpinned).
    Shapes: below / at / across the kernel's 2048-point tile and its 256-thread rounds, odd widths."""
    b = 6
    d = pf.synth.make_objects(b, h, w, seed=95, align_x0=1 if w % 4 else 4)
    d['mask'][4] = 0
    d['mask'][4, 5:9, 4:14] = 1                           # 40 points: below the 100-point rule -> untouched
    t = {k: d[k].cuda() for k in ('noc', 'depth', 'mask', 'bbox_xy0')}
    m1 = pf.statistical_outlier_mask(None, t['depth'], t['mask'], t['bbox_xy0'], source='depth')
    m2 = pf.statistical_outlier_mask(t['noc'], t['depth'], m1, t['bbox_xy0'], source='noc')
    for i in range(b):
        x0, y0 = (int(v) for v in d['bbox_xy0'][i])
        fd = np.zeros((240, 320), dtype=np.float32)
        fm = np.zeros((240, 320), dtype=bool)
        fd[y0:y0 + h, x0:x0 + w] = d['depth'][i].numpy()
        fm[y0:y0 + h, x0:x0 + w] = d['mask'][i].numpy() != 0
        noc_pts, depth_pts, (rows, cols) = po.crop_correspondences(np.transpose(d['noc'][i].numpy(), (1, 2, 0)), fd, fm,
                                                                   (x0, y0, x0 + w, y0 + h))
        k1 = po.statistical_outlier_indices(depth_pts) if depth_pts.shape[0] > 100 else np.arange(depth_pts.shape[0])
        want1 = np.zeros((h, w), dtype=np.uint8)
        want1[rows[k1] - y0, cols[k1] - x0] = 1
        np.testing.assert_array_equal(m1[i].cpu().numpy(), want1)
        k2 = po.run_pose_filters(noc_pts, depth_pts)
        want2 = np.zeros((h, w), dtype=np.uint8)
        want2[rows[k2] - y0, cols[k2] - x0] = 1
        np.testing.assert_array_equal(m2[i].cpu().numpy(), want2)
    n4 = int(((d['mask'][4] != 0) & (d['depth'][4] > 0)).sum())
    assert 30 <= n4 <= 40 and int(m1[4].sum()) == n4     # fewer than 100 points: left untouched (:311)
    if h * w > 1000:
        assert int(m1[0].sum()) < int(d['n_valid'][0])    # the gross outliers are gone
