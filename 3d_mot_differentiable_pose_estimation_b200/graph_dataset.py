"""Tracker-side consumers of the pose tensors, same names as the reference's
Tracking/datasets/graph_dataset.py:

  GraphDataset(rotations, translations, scales, input, instances_count, num_images).get_edge_data(...)
  GraphDataset(...).get_edge_data_office(...)

and the per-frame record of Detection/inference_detector.py:352-371 (`frame_records`).  The edge
construction runs on the GPU in three small launches for any number of sequences (`edge_features`);
the reference builds the same tensors with a quadruple Python loop and one torch op per pair.

Ground-truth association (train_utils.check_pair: 3-D box IoU against the frame's GT boxes) is a
labelling step of the tracker's training set, not part of this path: callers pass its per-node result
as `node_id` (-1 for None).  Without it every candidate pair is kept, which is exactly
get_edge_data_office.  relative_appearance (needs the voxel features) is not produced.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import torch

from . import _lib
from .function import _ptr, _stream


class EdgeBatch(NamedTuple):
    edge_index: torch.Tensor     # [2,E] int64, node indices local to each sequence, reference loop order
    edge_attr: torch.Tensor      # [E, 7+scale_dim] f32: dt(3), deuler(3), log scale ratio, frame distance
    targets: torch.Tensor        # [E] f32 (zeros without node_id)
    consecutive: torch.Tensor    # [E] int8
    edge_seq: torch.Tensor       # [E] int32 sequence of every edge
    false_positives: int
    seq_offsets: Optional[torch.Tensor] = None


def _frame_start(instances_count, n_sequences, n_frames, device):
    cnt = torch.as_tensor(instances_count, dtype=torch.int64).reshape(-1)
    if cnt.numel() != n_sequences * n_frames:
        raise ValueError(f'instances_count has {cnt.numel()} entries, expected {n_sequences}x{n_frames}')
    start = torch.zeros(cnt.numel() + 1, dtype=torch.int64)
    start[1:] = torch.cumsum(cnt, 0)
    return start.to(torch.int32).to(device), int(start[-1]), cnt


def max_candidate_edges(instances_count, n_sequences: int, n_frames: int, max_frame_dist: int, max_seq_len: int) -> int:
    """Upper bound on the number of directed edges (every candidate pair kept), computed on the host
    from the per-frame instance counts the caller already has (graph_dataset.py:24)."""
    cnt = torch.as_tensor(instances_count, dtype=torch.int64).reshape(n_sequences, n_frames)
    max_len = min(max_seq_len, n_frames)
    total = 0
    for d in range(1, max_frame_dist + 1):
        hi = max_len - d
        if hi <= 0:
            break
        lim = min(hi, n_frames - 1)
        total += int((cnt[:, :lim] * cnt[:, d:d + lim]).sum())
    return total


def edge_features(translations, rotations, scales, instances_count, n_sequences: int = 1,
                  n_frames: Optional[int] = None, node_id=None, max_frame_dist: int = 1, max_seq_len: int = 125) -> EdgeBatch:
    """All sequences of a batch in one call.  translations / rotations [N,3], scales [N] or [N,k] on a
    CUDA device (any float dtype; float64 is used as is), nodes ordered by (sequence, frame);
    instances_count: S*F per-frame counts (host)."""
    lib = _lib.lib()
    if not translations.is_cuda:
        raise _lib.PoseFitError('edge_features needs CUDA tensors: there is no CPU path')
    dev = translations.device
    cnt_flat = torch.as_tensor(instances_count).reshape(-1)
    if n_frames is None:
        n_frames = cnt_flat.numel() // n_sequences
    frame_start, n_nodes, _ = _frame_start(instances_count, n_sequences, n_frames, dev)
    t = translations.detach().to(torch.float64).reshape(-1, 3).contiguous()
    r = rotations.detach().to(torch.float64).reshape(-1, 3).contiguous()
    sc = scales.detach().to(torch.float64)
    sc = sc.reshape(sc.shape[0], -1).contiguous() if sc.dim() > 0 and sc.shape[0] > 0 else sc.reshape(0, 1)
    if t.shape[0] != n_nodes or r.shape[0] != n_nodes or sc.shape[0] != n_nodes:
        raise ValueError(f'{n_nodes} nodes by instances_count, got {t.shape[0]} / {r.shape[0]} / {sc.shape[0]} rows')
    scale_dim = int(sc.shape[1])
    nid = None
    if node_id is not None:
        nid = torch.as_tensor(node_id).to(device=dev, dtype=torch.int32).contiguous()
    e_max = max_candidate_edges(instances_count, n_sequences, n_frames, max_frame_dist, max_seq_len)
    a = 7 + scale_dim
    edge_index = torch.empty(2, max(e_max, 1), dtype=torch.int64, device=dev)
    edge_attr = torch.empty(max(e_max, 1), a, dtype=torch.float32, device=dev)
    targets = torch.zeros(max(e_max, 1), dtype=torch.float32, device=dev)
    consecutive = torch.empty(max(e_max, 1), dtype=torch.int8, device=dev)
    edge_seq = torch.empty(max(e_max, 1), dtype=torch.int32, device=dev)
    totals = torch.zeros(2, dtype=torch.int64, device=dev)
    ws_bytes = lib.posefit_edge_workspace_bytes(n_sequences, n_frames, n_nodes, max_frame_dist)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        code = lib.posefit_edge_features(_ptr(t), _ptr(r), _ptr(sc), scale_dim, _ptr(frame_start), _ptr(nid),
                                         n_sequences, n_frames, n_nodes, max_frame_dist, max_seq_len, e_max,
                                         _ptr(edge_index), _ptr(edge_attr), _ptr(targets), _ptr(consecutive),
                                         _ptr(edge_seq), _ptr(totals), _ptr(ws), ws_bytes, _stream(dev))
    _lib.check(code, 'posefit_edge_features')
    if node_id is None:
        n_edges, fp = e_max, 0            # every candidate is kept: no device read needed
    else:
        n_edges, fp = (int(v) for v in totals.cpu())
    return EdgeBatch(edge_index[:, :n_edges], edge_attr[:n_edges], targets[:n_edges], consecutive[:n_edges],
                     edge_seq[:n_edges], fp)


class GraphDataset:
    """Same constructor and get_edge_data signatures as Tracking/datasets/graph_dataset.py:10-29.
    `input` may carry the reference's per-frame dicts; only an optional 'node_id' list per frame (or
    the `node_id` argument) is used here -- see the module docstring."""

    def __init__(self, rotations, translations, scales, input, instances_count, num_images=25, appearance=None,
                 node_id=None):
        self.rotations = rotations
        self.translations = translations
        self.scales = scales
        self.input = input
        self.instances_count = [int(c) for c in instances_count]
        self.num_images = num_images
        self.appearance = appearance
        self.device = translations.device
        if node_id is None and input is not None and len(input) and isinstance(input[0], dict) and 'node_id' in input[0]:
            node_id = torch.cat([torch.as_tensor(f['node_id']).reshape(-1) for f in input[:num_images]])
        self.node_id = node_id

    def _edges(self, node_id, is_undirected, max_frame_dist, max_seq_len):
        eb = edge_features(self.translations, self.rotations, self.scales, self.instances_count[:self.num_images], 1,
                           self.num_images, node_id, max_frame_dist, max_seq_len)
        if eb.edge_index.shape[1] == 0:                                              # graph_dataset.py:180-183
            e = torch.tensor([], device=self.device)
            return e, e, e, e, None, 0, None
        edge_index, edge_attr, targets = eb.edge_index, eb.edge_attr, eb.targets
        if is_undirected:                                                            # :203-206
            edge_index = torch.cat((edge_index, torch.stack((edge_index[1], edge_index[0]))), dim=1)
            edge_attr = torch.cat((edge_attr, edge_attr), dim=0)
            targets = targets.repeat(2)
        return edge_index, edge_attr, targets, eb.consecutive, None, eb.false_positives, None

    def get_edge_data(self, is_undirected=True, max_frame_dist=1, max_seq_len=125, mode=None, vis_pose=False):
        if self.node_id is None:
            raise ValueError('get_edge_data needs the per-node ground-truth ids (check_pair results) as node_id; '
                             'use get_edge_data_office for the unlabelled graph')
        return self._edges(self.node_id, is_undirected, max_frame_dist, max_seq_len)

    def get_edge_data_office(self, is_undirected=True, max_frame_dist=1, max_seq_len=500):
        out = self._edges(None, is_undirected, max_frame_dist, max_seq_len)
        return out[0], out[1], out[3], out[4], out[6]


def frame_records(epilogue, status: torch.Tensor, frame_of: torch.Tensor, n_frames: int):
    """Per-frame records with the dataset names and shapes of the hdf5 files the detector writes for the
    tracker (Detection/inference_detector.py:352-371, read back by Tracking/datasets/front_dataset.py:74-80):
    'rotations' [n,3] XYZ Euler, 'translations' [n,3], 'scales' [n], 'pred_3Dbbox' [n,8,3], built from the
    batched `pose_epilogue` output (a PoseEpilogue); instances with status != 0 are dropped like the
    reference's rm_indicies (:207-209).  Returns (list of per-frame dicts of CUDA tensors, instances_count)."""
    keep = status == 0
    fo = frame_of.to(status.device)[keep].to(torch.int64)
    order = torch.argsort(fo, stable=True)
    fo = fo[order]
    euler, trans = epilogue.euler[keep][order], epilogue.global_trans[keep][order]
    scale, box = epilogue.global_scale[keep][order], epilogue.world_box[keep][order]
    counts = torch.bincount(fo, minlength=n_frames)[:n_frames]
    starts = [0] + torch.cumsum(counts, 0).cpu().tolist()
    frames = []
    for f in range(n_frames):
        a, b = starts[f], starts[f + 1]
        frames.append({'rotations': euler[a:b], 'translations': trans[a:b], 'scales': scale[a:b],
                       'pred_3Dbbox': box[a:b]})
    return frames, [starts[f + 1] - starts[f] for f in range(n_frames)]
