"""Pins the NumPy oracle (oracle/posefit_oracle.py) to outputs of the REAL reference
(tests/golden/*.npz, written by oracle/gen_golden.py from /root/reference)."""
import os

import numpy as np
import pytest

from oracle import posefit_oracle as po

TOL = 1e-11


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_umeyama_matches_reference(golden_dir):
    g = _load(golden_dir, 'umeyama.npz')
    names = [str(n) for n in g['names']]
    assert {'reflection', 'planar', 'zero_variance', 'single_point', 'repeated'} <= set(names)
    for k, name in enumerate(names):
        scales, rot_t, trans, tf = po.umeyama_fit(g[f'src_{k}'], g[f'dst_{k}'])
        np.testing.assert_allclose(scales, g[f'scales_{k}'], rtol=TOL, atol=TOL, err_msg=name)
        np.testing.assert_allclose(rot_t, g[f'rotation_{k}'], rtol=0, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(trans, g[f'translation_{k}'], rtol=TOL, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(tf, g[f'transform_{k}'], rtol=TOL, atol=1e-9, err_msg=name)


def test_zero_variance_scale_is_one(golden_dir):
    g = _load(golden_dir, 'umeyama.npz')
    k = [str(n) for n in g['names']].index('zero_variance')
    assert np.all(g[f'scales_{k}'] == 1.0)            # pose_utils.py:47-50
    assert np.all(po.umeyama_fit(g[f'src_{k}'], g[f'dst_{k}'])[0] == 1.0)
    np.testing.assert_array_equal(g[f'rotation_{k}'], np.identity(3))


def test_score_model_matches_reference(golden_dir):
    g = _load(golden_dir, 'evaluate.npz')
    for k in range(int(g['count'])):
        res, ratio, idx, _ = po.score_model(g[f'transform_{k}'], g[f'src_{k}'], g[f'dst_{k}'], float(g[f'pass_{k}']))
        np.testing.assert_allclose(res, g[f'residual_{k}'], rtol=1e-13)
        assert ratio == float(g[f'ratio_{k}'])
        np.testing.assert_array_equal(idx, g[f'idx_{k}'])


def test_index_zero_is_never_counted():
    # count_nonzero over the index array skips point 0 (pose_utils.py:10-12)
    src = np.zeros((4, 3))
    dst = np.zeros((4, 3))
    res, ratio, idx, _ = po.score_model(np.identity(4), src, dst, 1.0)
    assert list(idx) == [0, 1, 2, 3] and ratio == 3 / 4


def test_ransac_inlier_sets_match_reference(golden_dir):
    g = _load(golden_dir, 'ransac_inliers.npz')
    for k in range(int(g['count'])):
        src, dst = g[f'src_{k}'], g[f'dst_{k}']
        pass_t = float(g[f'pass_{k}'])
        rr = po.ransac_inliers(src, dst, g[f'idx_{k}'], pass_t, pass_t / 100)
        np.testing.assert_array_equal(src[rr['inlier_idx']], g[f'src_in_{k}'])
        np.testing.assert_array_equal(dst[rr['inlier_idx']], g[f'dst_in_{k}'])
        assert rr['ratio'] == float(g[f'ratio_{k}'])


def test_similarity_transform_matches_reference(golden_dir):
    g = _load(golden_dir, 'ransac.npz')
    seen = set()
    for k in range(int(g['count'])):
        name = str(g[f'name_{k}'])
        seen.add(name)
        out = po.similarity_transform(g[f'src_{k}'], g[f'dst_{k}'], g[f'idx_{k}'])
        assert out['ok'] == bool(g[f'ok_{k}']), name
        # number of hypotheses the reference actually drew (early stop, pose_utils.py:80-81)
        assert np.count_nonzero(~np.isnan(out['residuals'])) == int(g[f'calls_{k}']), name
        if out['ok']:
            np.testing.assert_allclose(out['scales'], g[f'scales_{k}'], rtol=TOL, err_msg=name)
            np.testing.assert_allclose(out['rot_t'], g[f'rotation_{k}'], atol=1e-9, err_msg=name)
            np.testing.assert_allclose(out['trans'], g[f'translation_{k}'], rtol=TOL, atol=1e-9, err_msg=name)
    assert 'identity_rot_early_stop' in seen and 'mostly_outliers' in seen
    k = [str(g[f'name_{i}']) for i in range(int(g['count']))].index('identity_rot_early_stop')
    assert int(g[f'calls_{k}']) < g[f'idx_{k}'].shape[0]          # the early stop really fired
    k = [str(g[f'name_{i}']) for i in range(int(g['count']))].index('mostly_outliers')
    assert not bool(g[f'ok_{k}'])                                 # the ratio<0.1 gate really fired


@pytest.mark.parametrize('tag,h,w,b', [('c1', 64, 64, 8), ('small', 24, 32, 6), ('odd', 19, 27, 4)])
def test_frames_match_reference(golden_dir, tag, h, w, b):
    g = _load(golden_dir, 'frames.npz')
    noc, depth, mask = g[f'{tag}_noc'], g[f'{tag}_depth'], g[f'{tag}_mask']
    xy0, idx, k_mat = g[f'{tag}_bbox_xy0'], g[f'{tag}_sample_idx'], g[f'{tag}_K']
    assert noc.shape == (b, 3, h, w)
    plain = po.batch_pose(noc, depth, mask, xy0, intrinsics=k_mat)
    n_valid = [int(g[f'{tag}_{i}_n_valid']) for i in range(b)]
    clamped = [np.minimum(idx[i], max(n_valid[i] - 1, 0)) for i in range(b)]
    rans = po.batch_pose(noc, depth, mask, xy0, intrinsics=k_mat, sample_idx=clamped)
    for i in range(b):
        status = int(g[f'{tag}_{i}_status'])
        assert plain[i]['n_valid'] == n_valid[i]
        assert rans[i]['status'] == status
        if status == 1:
            assert plain[i]['status'] == 1
            continue
        np.testing.assert_allclose(plain[i]['s'], g[f'{tag}_{i}_fit_scales'][0], rtol=TOL)
        np.testing.assert_allclose(plain[i]['rot_t'], g[f'{tag}_{i}_fit_rotation'], atol=1e-9)
        np.testing.assert_allclose(plain[i]['t'], g[f'{tag}_{i}_fit_translation'], rtol=TOL, atol=1e-9)
        ref_mask = np.unpackbits(g[f'{tag}_{i}_inlier_mask'])[:h * w].reshape(h, w)
        np.testing.assert_array_equal(rans[i]['inlier_mask'], ref_mask)
        assert rans[i]['ratio'] == float(g[f'{tag}_{i}_ratio'])
        np.testing.assert_allclose(rans[i]['pass_t'], float(g[f'{tag}_{i}_pass_t']), rtol=1e-14)
        if status == 0:
            np.testing.assert_allclose(rans[i]['s'], g[f'{tag}_{i}_ransac_scales'][0], rtol=TOL)
            np.testing.assert_allclose(rans[i]['rot_t'], g[f'{tag}_{i}_ransac_rotation'], atol=1e-9)
            np.testing.assert_allclose(rans[i]['t'], g[f'{tag}_{i}_ransac_translation'], rtol=TOL, atol=1e-9)
        if tag != 'c1':
            noc_pts, pts, _ = po.crop_correspondences(
                np.transpose(noc[i], (1, 2, 0)),
                _paste(depth[i], xy0[i], np.float32), _paste(mask[i] != 0, xy0[i], bool),
                (xy0[i, 0], xy0[i, 1], xy0[i, 0] + w, xy0[i, 1] + h), k_mat)
            np.testing.assert_array_equal(pts, g[f'{tag}_{i}_pts'])
            np.testing.assert_array_equal(noc_pts, g[f'{tag}_{i}_noc_pts'])
            if status == 0:
                cam = po.apply_similarity(g[f'{tag}_{i}_ransac_scales'], g[f'{tag}_{i}_ransac_rotation'],
                                          g[f'{tag}_{i}_ransac_translation'], noc_pts)
                np.testing.assert_allclose(cam, g[f'{tag}_{i}_transformed_pc'], rtol=1e-12, atol=1e-12)
                np.testing.assert_allclose(po.camera_to_world(cam, g[f'{tag}_{i}_campose']),
                                           g[f'{tag}_{i}_world_pc'], rtol=1e-12, atol=1e-12)


def _paste(crop, xy0, dtype):
    frame = np.zeros((po.FRAME_H, po.FRAME_W), dtype=dtype)
    h, w = crop.shape
    frame[int(xy0[1]):int(xy0[1]) + h, int(xy0[0]):int(xy0[0]) + w] = crop
    return frame


def test_clip_to_box_matches_reference(golden_dir):
    g = _load(golden_dir, 'clip.npz')
    sizes = []
    for k in range(int(g['count'])):
        keep = po.clip_to_box(g[f'pts_{k}'], g[f'box_{k}'], g[f'campose_{k}'])
        np.testing.assert_array_equal(keep, g[f'used_{k}'])
        np.testing.assert_array_equal(g[f'pts_{k}'][keep].reshape(-1, 3), g[f'new_depth_{k}'])
        sizes.append(len(keep))
    assert min(sizes) <= 20 < max(sizes)          # both sides of run_pose's "> 20" rule are covered


def test_sort_bbox_matches_reference(golden_dir):
    """sort_bbox (pose_estimation.py:72-93): the product's drop-in and the oracle's restatement against what the REAL
    function returned (oracle/gen_golden_bbox.py) -- axis-aligned boxes in the epilogue's corner order for every
    zero / positive extent pattern (argsort ties), shuffled corners, arbitrary and repeated points."""
    import importlib
    drop_in = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.pose_estimation')
    g = np.load(os.path.join(golden_dir, 'sort_bbox.npz'))
    n = int(g['n'])
    assert n >= 100
    for k in range(n):
        got = drop_in.sort_bbox(g[f'in_{k}'].copy())
        np.testing.assert_array_equal(got, g[f'out_{k}'], err_msg=str(g[f'tag_{k}']))


def test_rng_stream_position_matches_reference(golden_dir):
    """tests/golden/rng_stream.npz (oracle/gen_golden_rng.py: the real estimateSimilarityTransform on a seeded, UNPATCHED
    global stream): the oracle, fed with the draws made up front from the same seed, reproduces the outputs and the
    number of loop iterations; and rewinding to the seed and drawing iterations x 10 leaves np.random exactly where
    the reference's lazy per-iteration draws left it (the next four values of the stream)."""
    g = np.load(os.path.join(golden_dir, 'rng_stream.npz'))
    for name in [str(n) for n in g['names']]:
        src, dst = g[name + '_src'], g[name + '_dst']
        seed, ra, iters = int(g[name + '_seed']), float(g[name + '_ratio_adapt']), int(g[name + '_iterations'])
        np.random.seed(seed)
        idx = np.random.randint(src.shape[0], size=(100, 10))
        res = po.similarity_transform(src, dst, idx, ra)
        assert res['ok'] == bool(g[name + '_ok']), name
        assert res['iterations'] == iters, (name, res['iterations'], iters)
        np.testing.assert_allclose(res['out_transform'], g[name + '_transform'], rtol=1e-9, atol=1e-11, err_msg=name)
        np.random.seed(seed)
        if iters:
            np.random.randint(src.shape[0], size=(iters, 10))
        assert np.array_equal(np.random.randint(2 ** 31 - 1, size=4), g[name + '_next']), name
