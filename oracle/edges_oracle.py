"""TEST INFRASTRUCTURE -- CPU restatement of the tracker's edge construction,
GraphDataset.get_edge_data (Tracking/datasets/graph_dataset.py:30-199) and get_edge_data_office
(:231-342), with the ground-truth association (train_utils.check_pair) replaced by its per-node
result `node_id` (None / negative = no GT box matched).  Pinned to the real class by
oracle/gen_golden_edges.py -> tests/golden/edges.npz.  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


def get_edge_data(rotations, translations, scales, instances_count, num_images, node_id=None,
                  is_undirected=True, max_frame_dist=1, max_seq_len=125):
    """Returns (edge_index [2,E] int64, edge_attr [E,7+k] f32, targets [E] f32, consecutive [E] int8,
    false_positives).  node_id: per-node id or None (keep every pair, the _office variant)."""
    rotations = torch.as_tensor(rotations)
    translations = torch.as_tensor(translations)
    scales = torch.as_tensor(scales)
    if scales.dim() == 1:
        scales = scales[:, None]
    ids = None if node_id is None else [None if (i is None or int(i) < 0) else int(i) for i in node_id]
    rel_s, rel_p, rel_r, rel_t, edge_idxs, targets, consec = [], [], [], [], [], [], []
    false_positives = 0
    for t in range(num_images - 1):                                               # :50
        window = [f for f in range(t, t + 1 + max_frame_dist)                     # :56-61
                  if f != t and f >= 0 and f < min(max_seq_len, num_images)]
        start = int(sum(instances_count[:t]))                                     # :63-64
        end = start + instances_count[t]
        for j, frame in enumerate(window):                                        # :66
            prior = int(sum(instances_count[:frame]))                             # :71-72
            consec_n = prior + instances_count[frame]
            for n in range(start, end):                                           # :74
                id1 = None if ids is None else ids[n]
                if ids is not None:
                    if id1 is None and j == 0:                                    # :93-95
                        false_positives += 1
                        continue
                    elif id1 is None:                                             # :96-97
                        continue
                for m in range(prior, consec_n):                                  # :113
                    id2 = None if ids is None else ids[m]
                    if ids is not None:
                        if t == num_images - 2 and n == end - 1 and id2 is None:  # :133-136
                            false_positives += 1
                        if id2 is None:                                           # :145-146
                            continue
                        targets.append(1 if id1 == id2 else 0)                    # :139-144
                    consec.append(1 if frame == t + 1 else 0)                     # :149-162
                    edge_idxs.append([n, m])                                      # :164
                    rel_s.append(torch.log(scales[m, :] / scales[n, :])[None])    # :166-168
                    rel_p.append((translations[m, :] - translations[n, :])[None])  # :169-170
                    rel_r.append((rotations[m, :] - rotations[n, :])[None])       # :171-172
                    rel_t.append(torch.tensor([[frame - t]], dtype=torch.int64))  # :173-175
    if not rel_s:
        return None
    edge_attr = torch.cat((torch.cat(rel_p), torch.cat(rel_r), torch.cat(rel_s), torch.cat(rel_t)),
                          dim=-1).to(torch.float32)                               # :187-199
    edge_index = torch.tensor(edge_idxs, dtype=torch.long).t().contiguous()
    tg = torch.tensor(targets, dtype=torch.float32) if ids is not None else torch.zeros(len(edge_idxs))
    cm = torch.tensor(consec, dtype=torch.int8)
    if is_undirected:                                                             # :203-206
        edge_index = torch.cat((edge_index, torch.stack((edge_index[1], edge_index[0]))), dim=1)
        edge_attr = torch.cat((edge_attr, edge_attr), dim=0)
        tg = tg.repeat(2)
    return edge_index.numpy(), edge_attr.numpy(), tg.numpy(), cm.numpy(), false_positives


def make_sequence(rng: np.random.Generator, num_images=25, max_inst=6, p_unmatched=0.15, empty_frames=()):
    """A synthetic tracked sequence: per-frame instance counts, float64 pose arrays and per-node ids."""
    counts = [0 if f in empty_frames else int(rng.integers(1, max_inst + 1)) for f in range(num_images)]
    n = int(sum(counts))
    rot = rng.uniform(-np.pi, np.pi, size=(n, 3))
    trans = rng.normal(size=(n, 3)) * 2.0
    scales = rng.uniform(0.3, 2.5, size=(n, 1))
    ids = []
    for c in counts:
        perm = rng.permutation(max_inst * 2)[:c]
        ids += [int(i) if rng.uniform() > p_unmatched else -1 for i in perm]
    return counts, rot, trans, scales, np.asarray(ids, dtype=np.int64)
