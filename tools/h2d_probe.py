#!/usr/bin/env python
"""Copy-only probe of the end-to-end ceiling: every rank copies the bytes one e2e step ships (pinned host -> device,
non_blocking, one copy stream) in a loop, all ranks at once; no kernel runs.  Prints per-rank and aggregate GB/s.

    python tools/h2d_probe.py                               # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

If the aggregate scales with N and bench.py's e2e does not, the loss is in the pipeline (copy stream, allocator,
record_stream); if the aggregate saturates, the host side of the box is the ceiling."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get('RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 490.0          # bytes of one e2e step (16384 objects x 29.9 KB)
n_bufs = 5                                                       # the step's tensors: head, depth, mask, boxes, grads
sizes = [int(mb * 1e6 * f) for f in (0.315, 0.548, 0.137)] + [131072, 851968]
host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
devb = [[torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes] for _ in range(2)]
stream = torch.cuda.Stream()
total = sum(sizes)


def run(k):
    with torch.cuda.stream(stream):
        for i in range(k):
            for h, t in zip(host, devb[i & 1]):
                t.copy_(h, non_blocking=True)


run(3)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k = 20
with torch.cuda.stream(stream):
    a.record(stream)
run(k)
with torch.cuda.stream(stream):
    b.record(stream)
torch.cuda.synchronize()
ms = a.elapsed_time(b) / k
gbs = torch.tensor([total / ms / 1e6], device=dev, dtype=torch.float64)
if world > 1:
    allg = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(allg, gbs)
    vals = [float(x) for x in allg]
else:
    vals = [float(gbs)]
if rank == 0:
    print('h2d probe: %d rank(s), %.0f MB per step per rank: per-rank GB/s %s, aggregate %.1f GB/s'
          % (world, total / 1e6, ' '.join('%.1f' % v for v in vals), sum(vals)))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
