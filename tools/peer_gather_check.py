#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/peer_gather_check.py: PeerPoseGather vs the NCCL all-gather (values + timing)."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
shard = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.shard')
local_rank = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
rank, world = dist.get_rank(), dist.get_world_size()
n = 125000
pg = shard.PeerPoseGather(n)
for it in range(3):
    local = torch.randn(n, 16, dtype=torch.float64, device='cuda') + rank * 100 + it
    want = shard.gather_poses(local)
    pg.start(local)
    got = pg.wait()
    torch.cuda.synchronize()
    assert torch.equal(got, want), (rank, it)
    dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize(); dist.barrier()
ev[0].record()
for _ in range(20):
    shard.gather_poses(local)
ev[1].record()
torch.cuda.synchronize(); dist.barrier()
ev[2].record()
for _ in range(20):
    pg.start(local)
    pg.wait()
ev[3].record()
torch.cuda.synchronize()
if rank == 0:
    print('peer gather ok: world %d, nccl %.1f us, peer-copy %.1f us per gather of %d x 128 B'
          % (world, ev[0].elapsed_time(ev[1]) * 50, ev[2].elapsed_time(ev[3]) * 50, n))
dist.barrier()
dist.destroy_process_group()
