#!/usr/bin/env python
"""bench.py -- object poses/sec (fwd+bwd) of the B200 pose solver, with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W           # our arm (CUDA, through the C ABI)
    python bench.py --impl reference ...                     # the reference algorithm on host cores

A "step" = one pass of the hot path over this rank's shard of synthetic MOTFront-shaped
objects: forward (fused mask compaction + back-projection + moments + 3x3 solve) and backward
(adjoint + NOC-gradient scatter), i.e. BASELINE.json's metric "object poses/sec (fwd+bwd)".
Workload = the per-GPU shard of BASELINE config 5 at 8 GPUs (625 sequences x 25 frames x 8
objects = 125,000 objects of 64x64 crops per GPU, sharded by sequence, weak scaling); config 2
(4096 x 64x64 forward), config 3 (RANSAC, 128 hypotheses) and config 4 (384 x 112x112 fwd+bwd)
are timed too and reported under "configs".  Inputs are resident in HBM for `value`; `e2e` runs
the public autograd API from pinned host buffers with the copies inside the timed region.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = '3d_mot_differentiable_pose_estimation_b200'

SEQ_PER_GPU, FRAMES_PER_SEQ, OBJ_PER_FRAME = 625, 25, 8        # config 5 shard at 8 GPUs
METRIC = 'object poses/sec (fwd+bwd)'
UNIT = 'objects/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--objects', type=int, default=0,
                    help='objects per GPU per step (default: weak = the config-5 shard at 8 GPUs, 125000; strong = '
                         '1,000,000 / N)')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: 125000 objects per GPU whatever N (default); strong: config 5 as written -- 1M objects in '
                         'total, 1M / N per GPU')
    ap.add_argument('--size', type=int, default=64)
    ap.add_argument('--e2e-objects', type=int, default=16384)
    ap.add_argument('--cpu-sample', type=int, default=0, help='objects in the CPU baseline sample (0 = auto)')
    ap.add_argument('--gather', default='final', choices=['final', 'step'],
                    help='N > 1: one NCCL all-gather of all pose records after the last step (default, the north star\'s '
                         '"final gather"), or an all-gather of every step\'s records beside its backward pass')
    ap.add_argument('--no-extra', action='store_true', help='skip the config 2/3/4 side measurements')
    ap.add_argument('--no-full', action='store_true', help='skip the 1M-object single-GPU config-5 measurement (N = 1)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle = NumPy restatement of the reference; the reference itself is NumPy)
# ----------------------------------------------------------------------------------------------
def _reference_modules():
    """(pose_utils, pose_estimation) of the UNMODIFIED reference -- /root/reference in the build container, the copy
    oracle/Makefile staged under oracle/_ref/ on the GPU box -- or None (then the NumPy port is timed)."""
    if os.environ.get('POSEFIT_CPU_PORT'):
        return None
    try:
        from oracle import ref_import
        if not ref_import.reference_available():
            return None
        return ref_import.load_reference()
    except Exception as e:      # noqa: BLE001
        print('bench: reference modules not importable (%r), timing the port' % (e,), file=sys.stderr)
        return None


def cpu_kind():
    return 'reference' if _reference_modules() is not None else 'port'


def _cpu_worker(job):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    import contextlib
    import io
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import grad_oracle
    from oracle import posefit_oracle as po
    noc, depth, mask, xy0, idx, with_bwd, core = job
    if core is not None and hasattr(os, 'sched_setaffinity'):
        try:
            os.sched_setaffinity(0, {core})
        except OSError:
            pass
    ref = _reference_modules()
    if ref is not None:
        from oracle import ref_import
        pu, pe = ref
        intr = po.motfront_intrinsics()
    t0 = time.perf_counter()
    n = noc.shape[0]
    for i in range(n):
        h, w = depth[i].shape
        x0, y0 = int(xy0[i, 0]), int(xy0[i, 1])
        if ref is not None:
            # the per-instance flow of run_pose (pose_estimation.py:256-290, :323, :363) on the reference's own functions
            depth_pad = np.zeros((po.FRAME_H, po.FRAME_W))                          # :260-262 (float64 pads)
            depth_pad[y0:y0 + h, x0:x0 + w] = depth[i]
            mask_pad = np.zeros((po.FRAME_H, po.FRAME_W), dtype=bool)
            mask_pad[y0:y0 + h, x0:x0 + w] = mask[i] != 0
            noc_pad = np.zeros((po.FRAME_H, po.FRAME_W, 3))                         # :265-267
            noc_pad[y0:y0 + h, x0:x0 + w, :] = np.transpose(noc[i], (1, 2, 0))
            depth_pts, idxs = pe.backproject(depth_pad, intr, mask_pad)             # :290
            noc_pts = noc_pad[idxs[0], idxs[1], :] - 0.5                            # :323
            ok, inl = depth_pts.shape[0] > 0, None
            if ok and idx is None:
                src = np.transpose(np.hstack([noc_pts, np.ones([noc_pts.shape[0], 1])]))
                dst = np.transpose(np.hstack([depth_pts, np.ones([noc_pts.shape[0], 1])]))
                pu.estimateSimilarityUmeyama(src, dst)                              # pose_utils.py:16-61
            elif ok:
                with ref_import.replay_randint(idx[i], pu), contextlib.redirect_stdout(io.StringIO()):
                    ok = pu.estimateSimilarityTransform(noc_pts, depth_pts)[0] is not None   # :86-117
            if with_bwd and ok:
                wts = np.ones(noc_pts.shape[0])
                grad_oracle.fit_gradients(torch.from_numpy(noc_pts), torch.from_numpy(depth_pts), torch.from_numpy(wts),
                                          1.0, torch.ones(3, 3, dtype=torch.float64), torch.ones(3, dtype=torch.float64))
            continue
        fd = np.zeros((po.FRAME_H, po.FRAME_W), dtype=np.float32)
        fm = np.zeros((po.FRAME_H, po.FRAME_W), dtype=bool)
        fd[y0:y0 + h, x0:x0 + w] = depth[i]
        fm[y0:y0 + h, x0:x0 + w] = mask[i] != 0
        noc_pts, depth_pts, _ = po.crop_correspondences(np.transpose(noc[i], (1, 2, 0)), fd, fm,
                                                        (x0, y0, x0 + w, y0 + h))
        out = po.pose_from_correspondences(noc_pts, depth_pts, None if idx is None else idx[i])
        if with_bwd and out['status'] == 0:
            wts = np.zeros(noc_pts.shape[0])
            wts[out['inlier_idx']] = 1.0
            grad_oracle.fit_gradients(torch.from_numpy(noc_pts), torch.from_numpy(depth_pts), torch.from_numpy(wts),
                                      1.0, torch.ones(3, 3, dtype=torch.float64), torch.ones(3, dtype=torch.float64))
    return n, time.perf_counter() - t0


def cpu_objects_per_s(sample, with_bwd=True, ransac=False, procs=None, pin=False):
    """Times the CPU path over `sample` objects spread over `procs` processes; returns (obj/s, procs).  pin: bind the
    (single) worker to one core with sched_setaffinity."""
    import multiprocessing as mp
    import numpy as np
    procs = procs or (os.cpu_count() or 1)
    noc, depth, mask = sample['noc'].numpy(), sample['depth'].numpy(), sample['mask'].numpy()
    xy0 = sample['bbox_xy0'].numpy()
    idx = sample['sample_idx'].numpy() if ransac else None
    n = noc.shape[0]
    procs = max(1, min(procs, n))
    cuts = np.linspace(0, n, procs + 1).astype(int)
    cores = sorted(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else [0]
    jobs = [(noc[a:b], depth[a:b], mask[a:b], xy0[a:b], None if idx is None else idx[a:b], with_bwd,
             cores[0] if (pin and procs == 1) else None)
            for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    ctx = mp.get_context('spawn')
    with ctx.Pool(len(jobs)) as pool:
        pool.map(_cpu_worker, [(noc[:1], depth[:1], mask[:1], xy0[:1], None if idx is None else idx[:1], with_bwd, None)]
                 * len(jobs))                                     # import / warm-up
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    return n / dt, len(jobs)


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def _subsample(sample, n):
    return {k: v[:n] for k, v in sample.items()}


def cpu_baseline(pf, size, n_sample):
    """Headline baseline: fwd+bwd on all host cores; plus one pinned core, and the forward-only / RANSAC
    variants of configs 2 and 3 (each a bounded sample, stated)."""
    kind = cpu_kind()
    sample = pf.synth.make_objects(n_sample, size, size, seed=9001, n_hyp=128)
    value, cores = cpu_objects_per_s(sample, with_bwd=True)
    n1 = max(8, n_sample // max(cores, 1))
    one, _ = cpu_objects_per_s(_subsample(sample, n1), with_bwd=True, procs=1, pin=True)
    fwd, _ = cpu_objects_per_s(sample, with_bwd=False)
    n_r = max(cores, min(n_sample, 48 * cores))
    rans, _ = cpu_objects_per_s(_subsample(sample, n_r), with_bwd=False, ransac=True)
    what = ('the UNMODIFIED reference functions (PoseEst backproject on the padded 240x320 arrays + NOC gather + '
            'estimateSimilarityUmeyama / estimateSimilarityTransform with replayed np.random.randint)' if kind == 'reference'
            else 'the NumPy port of the reference (oracle/posefit_oracle.py)')
    return {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'cpu_model': cpu_model(),
            'single_core_value': one, 'single_core_pinned': hasattr(os, 'sched_setaffinity'),
            'config2_fwd_plain_value': fwd, 'config3_ransac128_value': rans,
            'sample': f'{n_sample} objects of the same {size}x{size} workload, fwd = {what}, fp64; bwd = fp64 torch-autograd '
                      f'restatement (the reference has no backward); {cores} processes x 1 thread; single core '
                      f'(sched_setaffinity) on {n1} objects; config 2 (fwd only) on {n_sample}, config 3 (RANSAC 128 hyp, '
                      f'replayed indices) on {n_r} objects'}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.samples, self.stop, self.th = index, [], False, None

    # NVML clocks-event-reason bits (nvml.h)
    BITS = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

    def _run_nvml(self):
        """NVML in-process: ~1 ms per sample instead of ~100 ms per nvidia-smi call, so even a 0.4 s timed
        region is covered by tens of samples.  Returns False when NVML is not usable."""
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = self.index
            if vis:
                ids = [v for v in vis.split(',') if v.strip()]
                if self.index < len(ids) and ids[self.index].strip().isdigit():
                    phys = int(ids[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            reasons = None
            for name in ('nvmlDeviceGetCurrentClocksEventReasons', 'nvmlDeviceGetCurrentClocksThrottleReasons'):
                fn = getattr(pynvml, name, None)
                if fn is None:
                    continue
                try:
                    int(fn(h))
                    reasons = fn
                    break
                except Exception:
                    continue
            if reasons is None:
                return False
        except Exception as e:
            print('bench: NVML clock sampling unavailable (%r), using nvidia-smi' % (e,), file=sys.stderr)
            return False
        while not self.stop:
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = int(reasons(h))
                self.samples.append([str(sm), str(mx)] + ['Active' if r & self.BITS[k] else 'Not Active' for k in
                                                         ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                                          'sw_power_cap')])
            except Exception as e:
                if not getattr(self, 'warned', False):
                    self.warned = True
                    print('bench: NVML sample failed: %r' % (e,), file=sys.stderr)
            time.sleep(0.005)
        return True

    def _run(self):
        if self._run_nvml():
            return
        while not self.stop:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        import statistics
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[j] for s in self.samples if len(s) >= 6 for j in range(4) if s[2 + j] == 'Active'})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
TOTAL_OBJECTS = 1000000                                        # config 5


def objects_per_gpu(args, world):
    if args.objects > 0:
        return args.objects
    if args.scaling == 'strong':
        return TOTAL_OBJECTS // max(world, 1) // (FRAMES_PER_SEQ * OBJ_PER_FRAME) * (FRAMES_PER_SEQ * OBJ_PER_FRAME)
    return SEQ_PER_GPU * FRAMES_PER_SEQ * OBJ_PER_FRAME


def workload_name(n_obj, size):
    """config.workload, shared by both arms."""
    return (f'config-5 shard: {n_obj} objects/GPU ({n_obj // (FRAMES_PER_SEQ * OBJ_PER_FRAME)} sequences x {FRAMES_PER_SEQ} '
            f'frames x {OBJ_PER_FRAME} objects) x {size}x{size} NOC+depth+mask crops, plain Umeyama fit fwd + bwd, '
            f'sharded by sequence')


def config_dict(args, n_obj, size):
    """The `config` object -- identical keys and values in both arms (our CUDA path and the CPU reference arm)."""
    P = size * size
    return {'workload': workload_name(n_obj, size), 'objects_per_gpu': n_obj, 'crop': [size, size],
            'scaling': args.scaling,
            'l2': 'inputs per step (%.1f GB) exceed the 126 MB L2; no flush needed' % (n_obj * 17 * P / 1e9),
            'collective': 'NCCL all-gather of the 128-B pose records (N > 1: final or per step, see run.collective; none at N = 1)'}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def time_kernels(fn, steps, warmup, torch):
    """fn() -> list of (name, start_event, end_event) recorded on the current stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    acc = {}
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    recs = []
    t_all0.record()
    for _ in range(steps):
        recs.append(fn())
    t_all1.record()
    torch.cuda.synchronize()
    for rec in recs:
        for name, e0, e1 in rec:
            acc.setdefault(name, []).append(e0.elapsed_time(e1))
    return t_all0.elapsed_time(t_all1) / steps, {k: sum(v) / len(v) for k, v in acc.items()}


def run_ours(args):
    import torch
    import torch.distributed as dist
    pf = importlib.import_module(PKG)
    lib = pf._lib.lib()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_node = pf.shard.bind_host_to_gpu(local) if world > 1 else None   # before any pinned allocation
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    hbm_peak, peak_src = peaks()
    size, n_obj = args.size, objects_per_gpu(args, world)
    P = size * size

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- resident shard (sharded by sequence: rank r owns sequences [r*625, (r+1)*625)) ---------
    d = pf.synth.make_objects(n_obj, size, size, seed=5000 + rank, device=dev)
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    g_s = torch.randn(n_obj, device=dev, generator=gen)
    g_R = torch.randn(n_obj, 9, device=dev, generator=gen)
    g_t = torch.randn(n_obj, 3, device=dev, generator=gen)
    kinv = pf.default_kinv(dev)

    # The one collective (north star: "no collective except a final NCCL gather of poses").  --gather final (default):
    # the steps of the timed region repeat the same job (this rank's shard of the 1M objects); its result -- the pose
    # records of the shard -- is all-gathered ONCE with NCCL after the last step, INSIDE the timed region.  --gather step: every step all-gathers its 128-byte records -- copies over
    # NVLink peer memory by the copy engines (shard.PeerPoseGather: no SM is taken from the HBM-bound backward kernel it
    # runs beside); POSEFIT_BENCH_GATHER=sync|async selects the NCCL all_gather after / beside the backward pass, which is
    # also the fallback when symmetric memory is not available.
    gather_mode = 'none'
    if world > 1:
        gather_mode = 'final' if args.gather == 'final' else os.environ.get('POSEFIT_BENCH_GATHER', 'peer')
    kept_poses = []
    peer = None
    if gather_mode == 'peer':
        ok = torch.ones(1, device=dev)
        try:
            peer = pf.shard.PeerPoseGather(n_obj)
        except Exception as e:      # noqa: BLE001  (symmetric memory is optional; NCCL always works)
            print('bench: peer-memory gather unavailable (%s), using NCCL' % str(e).splitlines()[0], file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            peer, gather_mode = None, 'sync'

    def step():
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        raw = pf.pose_fit_raw(d['noc'], d['depth'], d['mask'], d['bbox_xy0'], kinv)
        e1.record()
        work = None
        if gather_mode == 'peer':
            peer.start(raw.pose)                                 # copy engines, beside the backward pass
        elif gather_mode == 'async':
            _, work = pf.shard.gather_poses(raw.pose, async_op=True)
        pf.pose_fit_backward_raw(d['noc'], d['depth'], d['mask'], None, d['bbox_xy0'], kinv, raw.ctx, raw.status,
                                 g_s, g_R, g_t)
        e2.record()
        if gather_mode == 'peer':
            peer.wait()
        elif work is not None:
            work.wait()
        elif gather_mode == 'sync':
            pf.shard.gather_poses(raw.pose)                      # NCCL, after the backward pass, on the same stream
        elif gather_mode == 'final':
            kept_poses[:] = [raw.pose]                           # this step's records stay on this GPU
        return [('fit_moments_kernel', e0, e1), ('fit_backward_kernel', e1, e2)]

    for attempt in range(2):
        ok = torch.ones(1, device=dev)
        try:
            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001
            if gather_mode != 'peer':
                raise
            print('bench: peer-memory gather failed in warm-up (%s), using NCCL' % str(e).splitlines()[0], file=sys.stderr)
            ok.zero_()
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 1.0:
            break
        peer, gather_mode = None, 'sync'
    if world > 1:
        dist.barrier()
    launches0 = lib.posefit_launch_count()
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        t0, t1 = ev(), ev()
        t0.record()
        kept_poses.clear()
        recs = [step() for _ in range(args.steps)]
        gathered = None
        if gather_mode == 'final':
            gathered = pf.shard.gather_poses(kept_poses[0])      # the final NCCL gather, inside the timed region
        t1.record()
        torch.cuda.synchronize()
    if gathered is not None:
        assert gathered.shape[0] == world * n_obj
        del gathered
    kept_poses.clear()
    if world > 1:
        dist.barrier()
    launches = lib.posefit_launch_count() - launches0
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    kern = {}
    for rec in recs:
        for name, a, b in rec:
            kern.setdefault(name, []).append(a.elapsed_time(b))
    kern = {k: sum(v) / len(v) for k, v in kern.items()}
    value = world * n_obj / (ms * 1e-3)

    # algorithmic bytes per launch: fwd 17 B/px + 64 B pose record (+ 256 B ctx we
    # actually write is not counted), bwd 29 B/px + 52 B upstream grads + 256 B ctx read
    bytes_fwd = n_obj * (17 * P + 64)
    bytes_bwd = n_obj * (29 * P + 52)
    kernels = {
        'fit_moments_kernel': {'ms': kern['fit_moments_kernel'], 'algorithmic_bytes': bytes_fwd,
                               'gbs': bytes_fwd / kern['fit_moments_kernel'] / 1e6,
                               'note': 'interval also contains the dependent fit_solve_kernel (3x3 solves)'},
        'fit_backward_kernel': {'ms': kern['fit_backward_kernel'], 'algorithmic_bytes': bytes_bwd,
                                'gbs': bytes_bwd / kern['fit_backward_kernel'] / 1e6,
                                'note': 'interval also contains fit_backward_coef_kernel (per-object adjoint)'},
    }
    traffic = {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            tj = json.load(f)
        for kname in kernels:
            if kname in tj and size == 64:
                traffic[kname] = tj[kname]['bytes_per_object'] * n_obj
                kernels[kname]['traffic'] = traffic[kname]
    except (OSError, ValueError):
        pass
    dom = max(kernels, key=lambda k: kernels[k]['ms'])
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kernels[dom]['gbs'], 'peak': hbm_peak, 'unit': 'GB/s',
                'frac': kernels[dom]['gbs'] / hbm_peak, 'traffic': traffic.get(dom),
                'traffic_source': 'ncu --set full dram bytes per object (profiles/ncu_traffic.json) x objects per launch',
                'peak_source': peak_src,
                'step_achieved': (bytes_fwd + bytes_bwd) / ms / 1e6,
                'step_frac': (bytes_fwd + bytes_bwd) / ms / 1e6 / hbm_peak, 'kernels': kernels}

    # ---- end to end through the public autograd API from pinned host memory ---------------------
    # What crosses PCIe per object is what the reference's postprocess_dets holds for an instance
    # (Detection/tracker/postprocess.py:131-152): the 3x28x28 NOC head output (nocs_head.py:232-235), the depth and
    # mask windows of its box, the box corner -- 29.9 KB instead of the 69.7 KB of a materialised NOC crop.  On the
    # device: resample_noc (the per-instance roi_align resize, differentiable) -> pose_fit -> backward down to the
    # head output; the float32 poses go back to the host.
    ne = min(args.e2e_objects, n_obj)
    head_dev = torch.nn.functional.adaptive_avg_pool2d(d['noc'][:ne], 28).contiguous()      # synthetic head outputs
    roi_dev = torch.tensor([[size, size]], dtype=torch.int32, device=dev).repeat(ne, 1)
    host = {'head': head_dev.cpu().pin_memory(), 'roi_hw': roi_dev.cpu().pin_memory()}
    host.update({k: d[k][:ne].cpu().pin_memory() for k in ('depth', 'bbox_xy0')})
    # instance masks cross PCIe as bits (the masks are boolean, postprocess.py:134-139): packed on the host when the
    # sample is staged, expanded on the device by posefit_unpack_mask inside the timed step
    host['mask_bits'] = pf.pack_mask(d['mask'][:ne].cpu()).pin_memory()
    mask_shape = tuple(d['mask'][:ne].shape)
    del head_dev, roi_dev
    hg = {'s': g_s[:ne].cpu().pin_memory(), 'R': g_R[:ne].reshape(ne, 3, 3).cpu().pin_memory(),
          't': g_t[:ne].cpu().pin_memory()}
    out_host = torch.empty(ne, 13, dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + sum(v.numel() * v.element_size() for v in hg.values())
    d2h = out_host.numel() * 4

    # Two-stage pipeline, the way a data loader feeds a training step: the inputs of step i+1 cross
    # PCIe on a copy stream while step i computes.  Every step still copies ALL of its inputs from
    # pinned host memory and reads its result back; two device buffer sets, guarded by events.
    copy_stream = torch.cuda.Stream()
    slots = [{'ready': torch.cuda.Event(), 'free': torch.cuda.Event()} for _ in range(2)]
    for sl in slots:
        sl['free'].record()
    state = {'i': 0}

    def e2e_upload(sl):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl['free'])                   # the step that last used this set is done
            for k in ('head', 'roi_hw', 'depth', 'mask_bits', 'bbox_xy0'):
                sl[k] = host[k].to(dev, non_blocking=True)
            sl['g'] = tuple(hg[k].to(dev, non_blocking=True) for k in ('s', 'R', 't'))
            sl['ready'].record()

    def e2e_step():
        cur = torch.cuda.current_stream()
        sl = slots[state['i'] % 2]
        if 'head' not in sl or state['i'] == 0:
            e2e_upload(sl)                                       # first step: nothing was prefetched
        nxt = slots[(state['i'] + 1) % 2]
        cur.wait_event(sl['ready'])
        e2e_upload(nxt)                                          # next step's inputs, behind this step's compute
        head = sl['head'].requires_grad_(True)
        for t in (sl['head'], sl['roi_hw'], sl['depth'], sl['mask_bits'], sl['bbox_xy0']) + sl['g']:
            t.record_stream(cur)
        gs, gR, gt = sl['g']
        noc = pf.resample_noc(head, sl['roi_hw'], size, size)
        mask = pf.unpack_mask(sl['mask_bits'], mask_shape)
        scale, rot, trans, _, _, _ = pf.pose_fit(noc, sl['depth'], mask, sl['bbox_xy0'])
        torch.autograd.backward((scale, rot, trans), (gs, gR, gt))
        out_host[:, 0].copy_(scale.detach(), non_blocking=True)
        out_host[:, 1:10].copy_(rot.detach().reshape(ne, 9), non_blocking=True)
        out_host[:, 10:13].copy_(trans.detach(), non_blocking=True)
        sl['free'].record(cur)
        state['i'] += 1
        return head.grad

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    k_e2e = max(3, min(args.steps, 20))
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(k_e2e):
        e2e_step()
    torch.cuda.current_stream().wait_stream(copy_stream)         # every upload issued in the region ends in it
    t1.record()
    torch.cuda.synchronize()
    e2e_ms = t0.elapsed_time(t1) / k_e2e
    if world > 1:
        tms = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms)
    e2e = {'value': world * ne / (e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'objects_per_step_per_gpu': ne, 'ms_per_step': e2e_ms, 'h2d_gbs': h2d / e2e_ms / 1e6,
           'inputs': 'per object: 3x28x28 NOC head output + depth window + bit-packed mask window + box corner + upstream '
                     'gradients (%.1f KB); unpack_mask -> resample_noc -> pose_fit -> backward to the head output on the '
                     'device' % (h2d / ne / 1e3)}
    del host, hg

    # ---- side measurements: configs 2, 3, 4 (single GPU, rank 0) --------------------------------
    configs = {}
    if rank == 0 and not args.no_extra:
        # Small configs are timed as CUDA-graph replays that ROTATE over independent input sets whose
        # combined footprint is many times the 126 MB L2 (C2/C3: 4 x 285 MB, C4: 8 x 100 MB), so every
        # replay streams its inputs from HBM and no flush kernel runs next to the timed region
        # (tools/latency_probe.py: a write-flush leaves dirty lines the timed kernels then pay to evict,
        # a write+read flush measured 6 us slower on C2 than either; rotation has neither artefact).
        def timed(fns, reps=40, rounds=5):
            """(ms per replay with the replays queued back to back, ms of one replay on an idle GPU); fns = one closure
            per input set (the C ABI is capturable; graph replay removes the Python/ctypes launch overhead that
            otherwise dominates these 50-400 us configs).  Back to back -- `reps` replays between two events, like the K
            steps of the headline -- is the number the roofline fraction uses: an isolated replay also pays the host's
            graph-launch latency between the first event and the first kernel (~10 us), which is no property of the
            kernels.  Median of `rounds` rounds / of `reps` isolated replays."""
            graphs = []
            for fn in fns:
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fn()
                    side.synchronize()
                    with torch.cuda.graph(graph, stream=side):
                        fn()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graphs.append(graph)
            tt = []
            for i in range(reps + len(graphs)):
                a, b = ev(), ev()
                a.record()
                graphs[i % len(graphs)].replay()
                b.record()
                torch.cuda.synchronize()
                if i >= len(graphs):
                    tt.append(a.elapsed_time(b))
            tt.sort()
            bb = []
            for _ in range(rounds):
                a, b = ev(), ev()
                graphs[-1].replay()                                  # (the first timed replay does not start on an idle GPU)
                a.record()
                for i in range(reps):
                    graphs[i % len(graphs)].replay()
                b.record()
                torch.cuda.synchronize()
                bb.append(a.elapsed_time(b) / reps)
            bb.sort()
            return bb[len(bb) // 2], tt[len(tt) // 2]

        l2_note = 'graph replays rotate over %d independent input sets (%.0f MB in total, L2 is 126 MB); no flush kernel'
        launch_note = 'cuda graph replays queued back to back (40 per timed interval); ms_isolated = one replay on an idle GPU'
        c2s = [pf.synth.make_objects(4096, 64, 64, seed=2000 + i, device=dev, n_hyp=128) for i in range(4)]
        ms2, iso2 = timed([lambda c=c: pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv) for c in c2s])
        b2 = 4096 * (17 * 4096 + 64)
        configs['C2 4096x64x64 fwd plain'] = {'ms': ms2, 'ms_isolated': iso2, 'objects_per_s': 4096 / ms2 * 1e3, 'gbs': b2 / ms2 / 1e6,
                                             'frac': b2 / ms2 / 1e6 / hbm_peak, 'launch': launch_note,
                                             'l2': l2_note % (4, 4 * b2 / 1e6)}
        ms3, iso3 = timed([lambda c=c: pf.pose_fit_raw(c['noc'], c['depth'], c['mask'], c['bbox_xy0'], kinv,
                                                 sample_idx=c['sample_idx']) for c in c2s])
        b3 = 4096 * (17 * 4096 + 64 + 128 * 10 * 4 + 4096)
        configs['C3 4096x64x64 RANSAC 128 hyp'] = {'ms': ms3, 'ms_isolated': iso3, 'objects_per_s': 4096 / ms3 * 1e3, 'gbs': b3 / ms3 / 1e6,
                                                  'frac': b3 / ms3 / 1e6 / hbm_peak, 'scorer': 'closed-form moments',
                                                  'launch': launch_note, 'l2': l2_note % (4, 4 * b3 / 1e6)}
        del c2s
        c4s = [pf.synth.make_objects(384, 112, 112, seed=4000 + i, device=dev) for i in range(8)]
        g4 = (torch.randn(384, device=dev), torch.randn(384, 9, device=dev), torch.randn(384, 3, device=dev))

        def c4_step(c4):
            raw = pf.pose_fit_raw(c4['noc'], c4['depth'], c4['mask'], c4['bbox_xy0'], kinv)
            pf.pose_fit_backward_raw(c4['noc'], c4['depth'], c4['mask'], None, c4['bbox_xy0'], kinv, raw.ctx,
                                     raw.status, *g4)
        ms4, iso4 = timed([lambda c=c: c4_step(c) for c in c4s])
        b4 = 384 * (46 * 112 * 112)
        configs['C4 384x112x112 fwd+bwd'] = {'ms': ms4, 'ms_isolated': iso4, 'objects_per_s': 384 / ms4 * 1e3, 'gbs': b4 / ms4 / 1e6,
                                            'frac': b4 / ms4 / 1e6 / hbm_peak, 'launch': launch_note,
                                            'l2': l2_note % (8, 8 * b4 / 1e6)}
        del c4s

        # ---- the PUBLIC autograd operator on the headline shard: PoseFit.apply + backward (same inputs, same upstream
        # gradients through a linear loss); everything on the device side is the library's kernels
        g_R33 = g_R.reshape(n_obj, 3, 3)

        def api_step():
            noc = d['noc'].requires_grad_(True)
            noc.grad = None
            scale, rot, trans, _, _, _ = pf.pose_fit(noc, d['depth'], d['mask'], d['bbox_xy0'])
            torch.autograd.backward((scale, rot, trans), (g_s, g_R33, g_t))
            return noc.grad
        for _ in range(3):
            api_step()
        torch.cuda.synchronize()
        k_api = max(3, min(args.steps, 20))
        a0, a1 = ev(), ev()
        a0.record()
        for _ in range(k_api):
            api_step()
        a1.record()
        torch.cuda.synchronize()
        ms_api = a0.elapsed_time(a1) / k_api
        d['noc'].requires_grad_(False)
        b_api = bytes_fwd + bytes_bwd
        configs['C5 shard via PoseFit.apply + backward'] = {
            'ms': ms_api, 'objects_per_s': n_obj / ms_api * 1e3, 'gbs': b_api / ms_api / 1e6,
            'frac': b_api / ms_api / 1e6 / hbm_peak, 'vs_raw_api': ms_api / ms,
            'note': 'public torch.autograd.Function (pose_fit, kinv=None) + autograd backward, eager launches; scale / R / t '
                    'are written as float32 by the solve kernel (no eager torch arithmetic on the batch)'}

        # ---- the same shard fed by the NOC HEAD OUTPUTS: the roi_align resize fused into the loaders of the
        # fit and of its backward pass (PoseFitHead) against the composition it replaces (resample_noc -> pose_fit)
        head_dev = torch.nn.functional.adaptive_avg_pool2d(d['noc'], 28).contiguous()
        roi_dev = torch.tensor([[size, size]], dtype=torch.int32, device=dev).repeat(n_obj, 1)

        def head_step(fused):
            head = head_dev.requires_grad_(True)
            head.grad = None
            if fused:
                scale, rot, trans, _, _ = pf.pose_fit_head(head, roi_dev, d['depth'], d['mask'], d['bbox_xy0'])
            else:
                noc = pf.resample_noc(head, roi_dev, size, size)
                scale, rot, trans, _, _, _ = pf.pose_fit(noc, d['depth'], d['mask'], d['bbox_xy0'])
            torch.autograd.backward((scale, rot, trans), (g_s, g_R33, g_t))
        ms_head = {}
        for fused in (True, False):
            for _ in range(2):
                head_step(fused)
            torch.cuda.synchronize()
            a0, a1 = ev(), ev()
            a0.record()
            for _ in range(5):
                head_step(fused)
            a1.record()
            torch.cuda.synchronize()
            ms_head[fused] = a0.elapsed_time(a1) / 5
        hb = 3 * 28 * 28 * 4
        b_head = n_obj * ((5 * P + hb + 64) + (5 * P + hb + 52 + hb))
        configs['C5 shard from head outputs (fused loaders)'] = {
            'ms': ms_head[True], 'objects_per_s': n_obj / ms_head[True] * 1e3, 'gbs': b_head / ms_head[True] / 1e6,
            'frac': b_head / ms_head[True] / 1e6 / hbm_peak, 'composed_ms': ms_head[False],
            'note': 'pose_fit_head + backward: 3x28x28 head outputs sampled inside the loaders (%.1f KB of algorithmic traffic '
                    'per object instead of %.1f KB for resample_noc -> pose_fit -> backward -> resample backward); these '
                    'kernels are bound by instruction issue (bilinear taps, shared-memory atomics), not by HBM'
                    % (b_head / n_obj / 1e3, (46 * P + 24 * P + 2 * hb) / 1e3)}
        del head_dev, roi_dev

        # ---- config 5 as written: 1,000,000 objects on ONE GPU (69.6 GB of inputs + 49 GB of NOC gradient)
        if world == 1 and not args.no_full and n_obj < TOTAL_OBJECTS:
            free_b, _ = torch.cuda.mem_get_info()
            need = TOTAL_OBJECTS * (29 * P + 512)
            if free_b > need * 1.05:
                d.clear()
                torch.cuda.empty_cache()
                big = pf.synth.make_objects(TOTAL_OBJECTS, size, size, seed=5100, device=dev)
                gb = torch.Generator(device=dev).manual_seed(78)
                bg = (torch.randn(TOTAL_OBJECTS, device=dev, generator=gb), torch.randn(TOTAL_OBJECTS, 9, device=dev, generator=gb),
                      torch.randn(TOTAL_OBJECTS, 3, device=dev, generator=gb))
                g_noc = torch.empty_like(big['noc'])

                def big_step():
                    raw = pf.pose_fit_raw(big['noc'], big['depth'], big['mask'], big['bbox_xy0'], kinv)
                    pf.pose_fit_backward_raw(big['noc'], big['depth'], big['mask'], None, big['bbox_xy0'], kinv, raw.ctx,
                                             raw.status, *bg, out=g_noc)
                for _ in range(2):
                    big_step()
                torch.cuda.synchronize()
                k_big = 5
                a0, a1 = ev(), ev()
                a0.record()
                for _ in range(k_big):
                    big_step()
                a1.record()
                torch.cuda.synchronize()
                ms_big = a0.elapsed_time(a1) / k_big
                f0, f1, f2 = ev(), ev(), ev()                     # one more step, forward and backward apart
                f0.record()
                raw_b = pf.pose_fit_raw(big['noc'], big['depth'], big['mask'], big['bbox_xy0'], kinv)
                f1.record()
                pf.pose_fit_backward_raw(big['noc'], big['depth'], big['mask'], None, big['bbox_xy0'], kinv, raw_b.ctx,
                                         raw_b.status, *bg, out=g_noc)
                f2.record()
                torch.cuda.synchronize()
                del raw_b
                b_big = TOTAL_OBJECTS * (17 * P + 64 + 29 * P + 52)
                configs['C5 1M @ N=1'] = {'ms': ms_big, 'objects_per_s': TOTAL_OBJECTS / ms_big * 1e3,
                                          'gbs': b_big / ms_big / 1e6, 'frac': b_big / ms_big / 1e6 / hbm_peak,
                                          'steps': k_big, 'fwd_ms': f0.elapsed_time(f1), 'bwd_ms': f1.elapsed_time(f2),
                                          'free_gb_before': free_b / 1e9,
                                          'note': 'BASELINE config 5 on one GPU: 1,000,000 objects fwd + bwd per step'}
                del big, bg, g_noc
                torch.cuda.empty_cache()
            else:
                configs['C5 1M @ N=1'] = {'skipped': 'needs %.0f GB of device memory, %.0f GB free' % (need / 1e9, free_b / 1e9)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.cpu_sample or 256 * (os.cpu_count() or 1)
        cpu = cpu_baseline(pf, size, n_cpu)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling,
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(args, n_obj, size),
            'run': {'collective': {'none': 'none',
                                   'final': 'one NCCL all_gather of the 128-B pose records (the job\'s result) after the last step, '
                                            'inside the timed region',
                                   'peer': 'all-gather of 128-B pose records per step over NVLink peer memory '
                                           '(symmetric memory, copy engines) beside the backward pass',
                                   'sync': 'NCCL all_gather of 128-B pose records per step, after the backward pass',
                                   'async': 'NCCL all_gather of 128-B pose records per step, beside the backward pass'
                                   }[gather_mode],
                    'host_numa_node': numa_node},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches),
            'clocks': clk.summary(), 'configs': configs,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores (all of them), on a bounded sample
    of the same workload per step.  kind = "reference": the unmodified PoseEst functions staged under oracle/_ref (or
    /root/reference); "port": oracle/posefit_oracle.py when they are not there."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', args.gpus or 1))
    pf_synth = importlib.import_module(f'{PKG}.synth')
    cores = os.cpu_count() or 1
    n = args.cpu_sample or 32 * cores
    n_obj = objects_per_gpu(args, world)
    kind = cpu_kind()
    sample = pf_synth.make_objects(n, args.size, args.size, seed=9001)
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, used = cpu_objects_per_s(sample, with_bwd=True)
        vals.append(v)
    value = sorted(vals)[len(vals) // 2]
    fwd = ('the unmodified PoseEst backproject + NOC gather + estimateSimilarityUmeyama (fp64)' if kind == 'reference'
           else 'oracle port (NumPy restatement of PoseEst backproject + estimateSimilarityUmeyama, fp64)')
    what = (f'{n} objects per step of the same {args.size}x{args.size} workload (bounded sample, same generator); fwd = {fwd}; '
            f'bwd = fp64 torch-autograd restatement (the reference has no backward); {used} processes x 1 thread')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': n / value * 1e3, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_dict(args, n_obj, args.size),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': used, 'kind': kind, 'cpu_model': cpu_model(),
                             'sample': what},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    _emit(line)


def _emit(line):
    """The ONE line of stdout.  Everything else a run writes to file descriptor 1 -- NCCL's own `NCCL version ...`
    banner at communicator creation, a library's stray printf -- was pointed at stderr by main()."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


_REAL_STDOUT = 1

if __name__ == '__main__':
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
