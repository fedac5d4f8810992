"""Golden vectors for the tracker edge construction, produced by the REAL reference class
(Tracking/datasets/graph_dataset.py GraphDataset.get_edge_data / get_edge_data_office) together with the
real train_utils.check_pair / compute_3d_iou.  Build container only (/root/reference); writes
tests/golden/edges.npz.  Modules the reference imports at file scope but never touches on this path
(mathutils, torch_geometric, open3d, h5py, ...) are replaced by empty stand-ins.

Ground truth is made so that check_pair's answer is known: a matched node gets its own GT box as its
predicted box (IoU 1 -> that object's id), an unmatched node a box far away from every GT box (IoU 0 ->
None).  The ids check_pair really returned are read back from the reference's own output (targets)
and stored next to the inputs.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import edges_oracle  # noqa: E402

REF = os.environ.get('POSEFIT_REFERENCE_ROOT', '/root/reference')


def load_graph_dataset():
    sys.path.insert(0, REF)
    for name in ('mathutils', 'torch_geometric', 'torch_geometric.data', 'open3d', 'h5py', 'cv2', 'trimesh', 'mcubes',
                 'seaborn', 'motmetrics', 'matplotlib', 'matplotlib.pyplot'):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['torch_geometric.data'].Data = object
    spec = importlib.util.spec_from_file_location('ref_graph_dataset', os.path.join(REF, 'Tracking', 'datasets', 'graph_dataset.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def box_at(center, size):
    """8 corners in the order compute_3d_iou expects (train_utils.py:83-103): corners 0-3 the top face
    counter-clockwise in (x, z), 4-7 the bottom face, corner 0 above corner 4."""
    cx, cy, cz = center
    sx, sy, sz = size
    top = [(cx - sx, cy + sy, cz - sz), (cx - sx, cy + sy, cz + sz), (cx + sx, cy + sy, cz + sz), (cx + sx, cy + sy, cz - sz)]
    bot = [(x, cy - sy, z) for (x, _, z) in top]
    return np.asarray(top + bot, dtype=np.float64)


def build_inputs(counts, ids):
    """Per-frame dicts as front_dataset.py:82-99 hands them to GraphDataset."""
    frames, k = [], 0
    for f, c in enumerate(counts):
        fid = ids[k:k + c]
        gt_ids = sorted(set(int(i) for i in fid if i >= 0)) or [999]
        gt_boxes = np.stack([box_at((3.0 * g, 0.0, 0.0), (0.5, 0.5, 0.5)) for g in gt_ids])
        pred = []
        for j, i in enumerate(fid):
            pred.append(box_at((3.0 * int(i), 0.0, 0.0), (0.5, 0.5, 0.5)) if i >= 0
                        else box_at((-50.0 - 3.0 * j, 0.0, 40.0), (0.5, 0.5, 0.5)))
        frames.append({'gt_3Dbbox': torch.tensor(gt_boxes), 'gt_object_id': torch.tensor(gt_ids),
                       'pred_3Dbbox': torch.tensor(np.stack(pred)) if c else torch.zeros(0, 8, 3, dtype=torch.float64),
                       'translations': torch.zeros(c, 3), 'classes': torch.zeros(c, dtype=torch.int)})
        k += c
    return frames


def main():
    mod = load_graph_dataset()
    out = {}
    cases = [dict(seed=1, num_images=25, max_inst=5, p_unmatched=0.15, dist=1, undirected=True, empty=()),
             dict(seed=2, num_images=25, max_inst=6, p_unmatched=0.2, dist=5, undirected=True, empty=(3, 11)),
             dict(seed=3, num_images=8, max_inst=4, p_unmatched=0.0, dist=2, undirected=False, empty=()),
             dict(seed=4, num_images=6, max_inst=3, p_unmatched=0.5, dist=3, undirected=True, empty=(4,)),
             dict(seed=5, num_images=12, max_inst=8, p_unmatched=0.1, dist=5, undirected=False, empty=(), max_seq_len=9)]
    for ci, c in enumerate(cases):
        rng = np.random.default_rng(100 + c['seed'])
        counts, rot, trans, scales, ids = edges_oracle.make_sequence(rng, c['num_images'], c['max_inst'], c['p_unmatched'], c['empty'])
        frames = build_inputs(counts, ids)
        ds = mod.GraphDataset(torch.tensor(rot), torch.tensor(trans), torch.tensor(scales), frames, counts, num_images=c['num_images'])
        ds.device = torch.device('cpu')
        msl = c.get('max_seq_len', 125)
        ei, ea, tg, cm, _, fp, _ = ds.get_edge_data(is_undirected=c['undirected'], max_frame_dist=c['dist'], max_seq_len=msl)
        pre = f'c{ci}_'
        out[pre + 'counts'] = np.asarray(counts); out[pre + 'rot'] = rot; out[pre + 'trans'] = trans
        out[pre + 'scales'] = scales; out[pre + 'ids'] = ids
        out[pre + 'cfg'] = np.asarray([c['num_images'], c['dist'], int(c['undirected']), msl])
        out[pre + 'edge_index'] = ei.numpy(); out[pre + 'edge_attr'] = ea.numpy(); out[pre + 'targets'] = tg.numpy()
        out[pre + 'consecutive'] = cm.numpy(); out[pre + 'false_positives'] = np.asarray(fp)
        # the unlabelled (office) graph of the same sequence
        frames_o = [dict(f, rotations=torch.zeros(len(f['translations']), 3), scales=torch.ones(len(f['translations'])),
                         voxels=torch.zeros(len(f['translations']), 1)) for f in frames]
        ds_o = mod.GraphDataset(torch.tensor(rot), torch.tensor(trans), torch.tensor(scales), frames_o, counts, num_images=c['num_images'])
        ds_o.device = torch.device('cpu')
        ds_o.cad2world_mat = lambda *a, **k: None          # visualisation only (needs mathutils)
        ds_o.box2minmax = lambda *a, **k: None
        oi, oa, ocm, _, _ = ds_o.get_edge_data_office(is_undirected=c['undirected'], max_frame_dist=c['dist'], max_seq_len=msl)
        out[pre + 'office_edge_index'] = oi.numpy(); out[pre + 'office_edge_attr'] = oa.numpy()
        out[pre + 'office_consecutive'] = ocm.numpy()
        print(pre, 'edges', ei.shape, 'fp', fp, 'office', oi.shape)
    out['n_cases'] = np.asarray(len(cases))
    path = os.path.join(ROOT, 'tests', 'golden', 'edges.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == '__main__':
    main()
