"""CPU tests: the C-ABI library loads and exports every symbol include/posefit.h declares (no
compute calls without a GPU), argument validation, host-side partitioning, generator shapes, and the
N>1 gather path on a world_size-2 gloo group."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg


def test_library_exports_every_declared_symbol():
    lib_mod = load_pkg('_lib')
    path = lib_mod.build()
    header = open(os.path.join(ROOT, 'include', 'posefit.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(posefit_[a-z_]+)\s*\(', header))
    assert {'posefit_forward', 'posefit_forward_ransac', 'posefit_backward', 'posefit_points_forward',
            'posefit_points_forward_ransac', 'posefit_compact', 'posefit_points_evaluate',
            'posefit_transform_points', 'posefit_epilogue', 'posefit_clip_mask', 'posefit_sor_mask',
            'posefit_sor_workspace_bytes', 'posefit_resample_noc', 'posefit_resample_noc_backward',
            'posefit_gather_crops', 'posefit_workspace_bytes', 'posefit_backward_workspace_bytes', 'posefit_version',
            'posefit_error_string', 'posefit_launch_count'} <= declared
    handle = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(handle, name), f'{name} is declared in include/posefit.h but not exported'
    assert set(lib_mod.SYMBOLS) == declared
    assert lib_mod.lib().posefit_version() == 1


def test_argument_validation_without_gpu():
    lib = load_pkg('_lib').lib()
    assert lib.posefit_forward(None, None, None, None, None, 0, 0, 64, 64, None, None, None, None, None, 0, None) == 0
    assert lib.posefit_forward(None, None, None, None, None, 0, 4, 64, 64, None, None, None, None, None, 0, None) == -1
    one = ctypes.c_void_p(16)       # never dereferenced: validation happens first
    assert lib.posefit_forward(one, one, one, one, one, 0, 4, 0, 64, one, one, one, one, one, 0, None) == -2
    assert b'NULL' in lib.posefit_error_string(-1)
    assert b'workspace' in lib.posefit_error_string(-3)
    assert lib.posefit_launch_count() == 0


def test_product_refuses_cpu_tensors():
    pf = load_pkg()
    d = pf.synth.make_objects(2, 16, 16, seed=1)
    with pytest.raises(pf._lib.PoseFitError, match='no CPU path'):
        pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'])
    with pytest.raises(pf._lib.PoseFitError, match='no CPU path'):
        pf.points_fit_raw(torch.zeros(1, 3, 8, dtype=torch.float64), torch.zeros(1, 3, 8, dtype=torch.float64))


def test_product_does_not_import_the_oracle():
    import subprocess
    import sys
    code = ("import importlib, sys; sys.path.insert(0, %r); importlib.import_module(%r); "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
            % (ROOT, '3d_mot_differentiable_pose_estimation_b200'))
    subprocess.check_call([sys.executable, '-c', code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, '3d_mot_differentiable_pose_estimation_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.h')):
                assert 'oracle' not in open(os.path.join(dirpath, f)).read().replace('oracle/', ''), f


def test_generator_shapes_and_determinism():
    pf = load_pkg()
    a = pf.synth.make_objects(5, 24, 32, seed=3, n_hyp=7)
    b = pf.synth.make_objects(5, 24, 32, seed=3, n_hyp=7)
    assert a['noc'].shape == (5, 3, 24, 32) and a['noc'].dtype == torch.float32
    assert a['depth'].shape == (5, 24, 32) and a['mask'].dtype == torch.uint8
    assert a['sample_idx'].shape == (5, 7, 10) and a['sample_idx'].dtype == torch.int32
    for k in a:
        assert torch.equal(a[k], b[k]), k
    nv = ((a['mask'] != 0) & (a['depth'] > 0)).flatten(1).sum(1)
    assert torch.equal(nv.to(torch.int32), a['n_valid'])
    assert int((a['sample_idx'] >= a['n_valid'][:, None, None]).sum()) == 0
    assert float(a['noc'].min()) >= 0.0 and float(a['noc'].max()) <= 1.0
    assert (a['bbox_xy0'][:, 0] + 32 <= 320).all() and (a['bbox_xy0'][:, 1] + 24 <= 240).all()


def test_sequence_shard_partitions_everything():
    shard = load_pkg('shard')
    for n, world in ((5000, 8), (625, 1), (7, 3), (3, 8)):
        cover = []
        for r in range(world):
            s, e = shard.sequence_shard(n, r, world)
            assert 0 <= s <= e <= n
            cover.extend(range(s, e))
        assert cover == list(range(n))
        sizes = [shard.sequence_shard(n, r, world)[1] - shard.sequence_shard(n, r, world)[0] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert shard.sequence_shard(5000, 3, 8) == (1875, 2500)        # BASELINE config 5: 625 sequences per GPU
    off = torch.arange(0, 11) * 200
    assert shard.object_range(off, 1, 2) == (1000, 2000)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    shard = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.shard')
    s0, s1 = shard.sequence_shard(10, rank, world)
    local = torch.arange(s0 * 4, s1 * 4, dtype=torch.float64)[:, None].repeat(1, 16) + 0.5
    eq = shard.gather_poses(local)
    eq_async, work = shard.gather_poses(local, async_op=True)      # the form bench.py overlaps with the backward pass
    work.wait()
    assert torch.equal(eq_async, eq)
    counts = [3, 5]
    ragged = shard.gather_poses(torch.full((counts[rank], 16), float(rank), dtype=torch.float64), counts=counts)
    q.put((rank, eq[:, 0].tolist(), ragged[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_poses_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, eq, ragged in got:
        assert eq == [i + 0.5 for i in range(40)]                  # rank order, every object exactly once
        assert ragged == [0.0] * 3 + [1.0] * 5


def test_numa_binding_helpers(tmp_path):
    """bind_host_to_gpu parses sysfs cpulists and is a no-op (None) when the platform gives no answer."""
    import importlib
    shard = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.shard')
    assert shard._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert shard._parse_cpulist('') == set()
    before = os.sched_getaffinity(0)
    assert shard.bind_host_to_gpu(0, sysfs=str(tmp_path)) is None   # no CUDA device / no sysfs entry here
    assert os.sched_getaffinity(0) == before


def test_bucket_groups_partition_the_instances():
    """run_pose_batched(bucket=k): every instance lands in exactly one group whose canvas holds its box, widths are
    multiples of 4, members stay in instance order (the RANSAC draws are made in that order)."""
    import importlib
    fe = importlib.import_module('3d_mot_differentiable_pose_estimation_b200.frontend')
    rng = np.random.default_rng(3)
    x0 = rng.integers(0, 200, size=40)
    y0 = rng.integers(0, 150, size=40)
    w = rng.integers(1, 120, size=40)
    h = rng.integers(1, 90, size=40)
    boxes = np.stack([x0, y0, x0 + w, y0 + h], axis=1).astype(np.int32)
    boxes[7, 2:] = boxes[7, :2]                                   # a degenerate (empty) box still gets a 1-pixel canvas
    for bucket in (1, 16, 32, 50):
        groups = fe.bucket_groups(torch.from_numpy(boxes), bucket)
        seen = sorted(i for members in groups.values() for i in members)
        assert seen == list(range(40))
        for (gh, gw), members in groups.items():
            assert members == sorted(members) and gw % 4 == 0 and gh >= 1
            for i in members:
                bh, bw = max(boxes[i, 3] - boxes[i, 1], 1), max(boxes[i, 2] - boxes[i, 0], 1)
                assert bh <= gh < bh + bucket and bw <= gw < bw + bucket + 4
    assert len(fe.bucket_groups(torch.from_numpy(boxes), 1000)) == 1
