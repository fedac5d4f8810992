"""Multi-GPU partitioning of the pose path.

Objects are independent, so the path shards with no data-path collective: sequences (25 frames,
Detection/train_combined.py:128-129, :237-246) are dealt out in contiguous blocks, every rank fits
the objects of its own sequences on its own GPU, and ONE all-gather of the 128-byte pose records
makes the result visible everywhere (the reference only ever gathers python prediction lists,
Detection/evaluator/FrontEvaluator.py:143).  Gradients stay with the rank that owns the NOC crop.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def sequence_shard(n_sequences: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, end) of the whole sequences owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    base, rem = divmod(n_sequences, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def object_range(seq_offsets: torch.Tensor, rank: int, world_size: int) -> Tuple[int, int]:
    """seq_offsets: [n_sequences + 1] prefix of objects per sequence (objects are stored sequence
    by sequence).  Returns the contiguous object range of this rank's sequences."""
    s0, s1 = sequence_shard(int(seq_offsets.numel()) - 1, rank, world_size)
    return int(seq_offsets[s0]), int(seq_offsets[s1])


def gather_poses(local: torch.Tensor, counts: Optional[list] = None, group=None, async_op: bool = False):
    """All-gather per-rank pose records [n_r, 16] into [sum n_r, 16] in rank order.
    counts: objects per rank when shards are ragged (None = equal shards).
    async_op (equal shards only): returns (out, work) right away -- the collective runs on the
    backend's own stream, ordered after the work already queued on the current stream, so the caller
    can queue the backward pass next to it and `work.wait()` when the gathered poses are needed."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return (local, None) if async_op else local
    world = dist.get_world_size(group)
    if counts is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        work = dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    if async_op:
        raise ValueError('async_op needs equal shards')
    if len(counts) != world or counts[dist.get_rank(group)] != local.shape[0]:
        raise ValueError('counts must list the shard size of every rank')
    width = max(counts)
    padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width:r * width + counts[r]] for r in range(world)], dim=0)


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int, sysfs: str = '/sys') -> Optional[int]:
    """Pin this process to the CPU cores of the NUMA node its GPU hangs off.  One process per GPU
    stages its crops through pinned host buffers; first-touching those buffers from local cores keeps
    the 8 concurrent host->device streams off the inter-socket link.  Returns the node, or None when
    the platform does not say (VMs report numa_node = -1) -- then nothing is changed."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(os.path.join(sysfs, 'bus/pci/devices', bus, 'numa_node')) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(os.path.join(sysfs, 'devices/system/node/node%d/cpulist' % node)) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


class PeerPoseGather:
    """All-gather of the pose records over NVLink peer memory, moved by the copy engines.

    Every rank owns TWO [world * n_local, width] buffers in symmetric memory (torch.distributed._symmetric_memory:
    each buffer is mapped into every peer's address space), used alternately.  `start(local)` queues, on a side
    stream and behind the work already on the current stream, one device-to-device copy of this rank's records into
    its slot of EVERY rank's buffer for this step, then a signal-pad barrier; `wait()` makes the current stream wait
    for that barrier and returns the gathered tensor.  No SM is occupied by the transfer, so unlike an NCCL all-gather
    it can run beside an HBM-bound kernel without displacing its CTAs.

    Why two buffers: with one, a fast rank's copies for step n+1 could land in a slower peer's buffer while that
    peer's kernels still read step n's records.  With two, step n+1 writes the OTHER buffer; step n+2 reuses step n's,
    and by then the barrier of step n+1 has passed -- every rank's step-(n+1) copies, which that rank's side stream
    ordered behind everything its compute stream had queued (its readers of step n included), are complete.  So the
    tensor returned by `wait()` stays valid until the caller's `start()` two steps later, provided its readers were
    queued on the current stream before the next `start()`.
    Construction is collective; it raises if symmetric memory is not available (callers fall back to
    `gather_poses`, the NCCL path)."""

    def __init__(self, n_local: int, width: int = 16, dtype=torch.float64, group=None, timeout_ms: int = 20000):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.n = int(n_local)
        self.timeout_ms = int(timeout_ms)
        dev = torch.device('cuda', torch.cuda.current_device())
        self.bufs, self.hdls, self.peers = [], [], []
        for _ in range(2):
            buf = symm.empty(self.world * self.n, width, dtype=dtype, device=dev)
            hdl = symm.rendezvous(buf, self.group)
            self.bufs.append(buf)
            self.hdls.append(hdl)
            self.peers.append([hdl.get_buffer(r, (self.world * self.n, width), dtype) for r in range(self.world)])
        self.buf = self.bufs[0]                                  # (shape / dtype reference)
        self.step = 0
        self.cur = 0
        self.stream = torch.cuda.Stream()
        self.done = torch.cuda.Event()

    def start(self, local: torch.Tensor) -> None:
        if tuple(local.shape) != (self.n, self.buf.shape[1]) or local.dtype != self.buf.dtype:
            raise ValueError('local records do not match the gather buffer')
        local = local.contiguous()
        self.cur = self.step & 1
        self.step += 1
        peers, hdl = self.peers[self.cur], self.hdls[self.cur]
        self.stream.wait_stream(torch.cuda.current_stream())
        lo = self.rank * self.n
        with torch.cuda.stream(self.stream):
            local.record_stream(self.stream)
            for k in range(self.world):
                r = (self.rank + k) % self.world                 # every rank starts with a different target
                peers[r][lo:lo + self.n].copy_(local, non_blocking=True)
            hdl.barrier(channel=0, timeout_ms=self.timeout_ms)   # all ranks' copies of this step have landed
            self.done.record(self.stream)

    def wait(self) -> torch.Tensor:
        torch.cuda.current_stream().wait_event(self.done)
        return self.bufs[self.cur]
