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
# Skewed ranks, the gathered tensor READ on every step by a slow consumer kernel that is still running when a faster
# peer has already started the next step's copies: with the two alternating buffers no step may see another step's records.
torch.cuda.synchronize(); dist.barrier()
checks = []
for it in range(12):
    local = torch.full((n, 16), float(1000 * it + rank), dtype=torch.float64, device='cuda')
    if it % world == rank:
        torch.cuda._sleep(20_000_000)                            # ~10 ms: this rank lags behind its peers
    pg.start(local)
    got = pg.wait()
    acc = torch.zeros(world, dtype=torch.float64, device='cuda')
    for rep in range(4 if (it + 1) % world == rank else 1):      # a slow reader of THIS step's records
        acc = acc + got.view(world, n * 16).double().mean(dim=1) / (4 if (it + 1) % world == rank else 1)
    checks.append((it, acc))
torch.cuda.synchronize()
for it, acc in checks:
    want = torch.tensor([1000.0 * it + r for r in range(world)], dtype=torch.float64, device='cuda')
    assert torch.allclose(acc, want, rtol=0, atol=1e-9), (rank, it, acc.tolist(), want.tolist())
dist.barrier()
if rank == 0:
    print('peer gather: 12 skewed steps read back correctly on every rank')
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
