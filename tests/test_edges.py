"""Tracker edge construction: oracle vs the golden vectors of the real
GraphDataset (CPU), CUDA kernels vs golden vectors and oracle (GPU)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import edges_oracle as eo  # noqa: E402

GOLD = np.load(os.path.join(ROOT, 'tests', 'golden', 'edges.npz'))
N_CASES = int(GOLD['n_cases'])


def _case(i):
    pre = f'c{i}_'
    g = {k[len(pre):]: GOLD[k] for k in GOLD.files if k.startswith(pre)}
    num_images, dist, undirected, msl = (int(v) for v in g['cfg'])
    return g, num_images, dist, bool(undirected), msl


@pytest.mark.parametrize('i', range(N_CASES))
def test_oracle_matches_reference_golden(i):
    g, num_images, dist, undirected, msl = _case(i)
    ei, ea, tg, cm, fp = eo.get_edge_data(g['rot'], g['trans'], g['scales'], g['counts'].tolist(), num_images, g['ids'],
                                          undirected, dist, msl)
    assert np.array_equal(ei, g['edge_index'])
    assert np.array_equal(ea, g['edge_attr'])                 # same torch ops: bit-exact
    assert np.array_equal(tg, g['targets'])
    assert np.array_equal(cm, g['consecutive'])
    assert fp == int(g['false_positives'])
    oi, oa, _, ocm, _ = eo.get_edge_data(g['rot'], g['trans'], g['scales'], g['counts'].tolist(), num_images, None,
                                         undirected, dist, msl)
    assert np.array_equal(oi, g['office_edge_index'])
    assert np.array_equal(oa, g['office_edge_attr'])
    assert np.array_equal(ocm, g['office_consecutive'])


def _check_attr(ours, ref):
    """Differences and the frame distance are exact (same fp64 subtraction, one rounding); the log
    column may differ by one float32 ulp (CUDA log vs the host libm before the rounding)."""
    ours, ref = np.asarray(ours), np.asarray(ref)
    assert ours.shape == ref.shape
    assert np.array_equal(ours[:, :6], ref[:, :6])
    assert np.array_equal(ours[:, -1], ref[:, -1])
    np.testing.assert_allclose(ours[:, 6:-1], ref[:, 6:-1], rtol=2.4e-7, atol=1e-45)


@pytest.mark.gpu
@pytest.mark.parametrize('i', range(N_CASES))
def test_gpu_graph_dataset_matches_reference_golden(i):
    pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
    from importlib import import_module
    gd = import_module('3d_mot_differentiable_pose_estimation_b200.graph_dataset')
    g, num_images, dist, undirected, msl = _case(i)
    dev = torch.device('cuda', 0)
    ds = gd.GraphDataset(torch.tensor(g['rot'], device=dev), torch.tensor(g['trans'], device=dev),
                         torch.tensor(g['scales'], device=dev), None, g['counts'].tolist(), num_images=num_images,
                         node_id=torch.tensor(g['ids']))
    ei, ea, tg, cm, _, fp, _ = ds.get_edge_data(is_undirected=undirected, max_frame_dist=dist, max_seq_len=msl)
    assert ei.dtype == torch.int64 and ea.dtype == torch.float32 and cm.dtype == torch.int8
    assert np.array_equal(ei.cpu().numpy(), g['edge_index'])
    _check_attr(ea.cpu().numpy(), g['edge_attr'])
    assert np.array_equal(tg.cpu().numpy(), g['targets'])
    assert np.array_equal(cm.cpu().numpy(), g['consecutive'])
    assert fp == int(g['false_positives'])
    oi, oa, ocm, _, _ = ds.get_edge_data_office(is_undirected=undirected, max_frame_dist=dist, max_seq_len=msl)
    assert np.array_equal(oi.cpu().numpy(), g['office_edge_index'])
    _check_attr(oa.cpu().numpy(), g['office_edge_attr'])
    assert np.array_equal(ocm.cpu().numpy(), g['office_consecutive'])
    assert pf._lib.lib().posefit_launch_count() > 0


@pytest.mark.gpu
def test_gpu_edge_features_batched_sequences():
    """Many sequences in one call == the oracle run sequence by sequence (incl. empty frames, all
    nodes unmatched, a sequence without any detection)."""
    gd = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.graph_dataset')
    rng = np.random.default_rng(7)
    dev = torch.device('cuda', 0)
    F, S, D = 25, 37, 5
    seqs = []
    for s in range(S):
        empty = (2, 3) if s % 5 == 0 else ()
        seq = eo.make_sequence(rng, F, 7, 0.2 if s != 4 else 1.0, empty)
        if s == 9:
            seq = ([0] * F, np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 1)), np.zeros(0, dtype=np.int64))
        seqs.append(seq)
    counts = sum((q[0] for q in seqs), [])
    rot = np.concatenate([q[1] for q in seqs]); trans = np.concatenate([q[2] for q in seqs])
    scales = np.concatenate([q[3] for q in seqs]); ids = np.concatenate([q[4] for q in seqs])
    for node_id in (ids, None):
        eb = gd.edge_features(torch.tensor(trans, device=dev), torch.tensor(rot, device=dev), torch.tensor(scales, device=dev),
                              counts, S, F, None if node_id is None else torch.tensor(node_id), D, 125)
        ei_all, ea_all, tg_all, cm_all, fp_all = [], [], [], [], 0
        seq_of = []
        for s, q in enumerate(seqs):
            r = eo.get_edge_data(q[1], q[2], q[3], q[0], F, None if node_id is None else q[4], False, D, 125)
            if r is None:
                continue
            ei_all.append(r[0]); ea_all.append(r[1]); tg_all.append(r[2]); cm_all.append(r[3]); fp_all += r[4]
            seq_of += [s] * r[0].shape[1]
        ei = np.concatenate(ei_all, axis=1)
        assert np.array_equal(eb.edge_index.cpu().numpy(), ei)
        _check_attr(eb.edge_attr.cpu().numpy(), np.concatenate(ea_all))
        assert np.array_equal(eb.consecutive.cpu().numpy(), np.concatenate(cm_all))
        assert np.array_equal(eb.edge_seq.cpu().numpy(), np.asarray(seq_of, dtype=np.int32))
        if node_id is not None:
            assert np.array_equal(eb.targets.cpu().numpy(), np.concatenate(tg_all))


@pytest.mark.gpu
def test_gpu_frame_records_feed_the_graph():
    """pose_epilogue -> frame_records -> GraphDataset: the tracker runs straight off GPU pose tensors."""
    pf = importlib.import_module('3d_mot_differentiable_pose_estimation_b200')
    gd = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.graph_dataset')
    dev = torch.device('cuda', 0)
    F, per = 5, 4
    d = pf.synth.make_objects(F * per, 32, 32, seed=11, device=dev)
    raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], pf.default_kinv(dev))
    epi = pf.pose_epilogue(raw, d['depth'], d['mask'], d['bbox_xy0'], pf.default_kinv(dev))
    frame_of = torch.arange(F * per, device=dev) // per
    frames, counts = gd.frame_records(epi, raw.status, frame_of, F)
    assert sum(counts) == int((raw.status == 0).sum())
    rot = torch.cat([f['rotations'] for f in frames]); tr = torch.cat([f['translations'] for f in frames])
    sc = torch.cat([f['scales'] for f in frames])
    assert frames[0]['pred_3Dbbox'].shape[1:] == (8, 3)
    ds = gd.GraphDataset(rot, tr, sc, frames, counts, num_images=F)
    ei, ea, cm, _, _ = ds.get_edge_data_office(is_undirected=True, max_frame_dist=2)
    ref = eo.get_edge_data(rot.cpu().numpy(), tr.cpu().numpy(), sc.cpu().numpy(), counts, F, None, True, 2, 500)
    assert np.array_equal(ei.cpu().numpy(), ref[0])
    _check_attr(ea.cpu().numpy(), ref[1])
