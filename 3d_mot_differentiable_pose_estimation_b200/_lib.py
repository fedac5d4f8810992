"""ctypes binding of the C-ABI library declared in include/posefit.h.

The CUDA library is the product: there is no Python or CPU fallback.  If
`libposefit_b200.so` is missing it is built in-tree with nvcc (csrc/Makefile); if that is
impossible the import fails loudly.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# POSEFIT_LIB: tooling only (A/B runs of experimental builds, tools/ransac_phases.py); the product is the in-tree library
LIB_PATH = os.environ.get('POSEFIT_LIB') or os.path.join(_HERE, 'libposefit_b200.so')
CSRC = os.path.join(_HERE, 'csrc')

POSE_DOUBLES = 16
CTX_DOUBLES = 32
ABI_VERSION = 1

# every symbol include/posefit.h declares
SYMBOLS = ('posefit_version', 'posefit_error_string', 'posefit_workspace_bytes', 'posefit_forward',
           'posefit_forward_ransac', 'posefit_backward', 'posefit_backward_workspace_bytes', 'posefit_launch_count',
           'posefit_points_forward', 'posefit_points_forward_ransac', 'posefit_compact',
           'posefit_points_evaluate', 'posefit_transform_points', 'posefit_epilogue', 'posefit_clip_mask', 'posefit_sor_mask',
           'posefit_sor_workspace_bytes', 'posefit_resample_noc', 'posefit_resample_noc_backward',
           'posefit_gather_crops', 'posefit_unpack_mask', 'posefit_edge_features', 'posefit_edge_workspace_bytes',
           'posefit_debug_reload_env', 'posefit_forward_ex', 'posefit_forward_ransac_ex',
           'posefit_forward_head', 'posefit_backward_head', 'posefit_head_workspace_bytes')

_lock = threading.Lock()
_lib = None


class PoseFitError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into libposefit_b200.so (no GPU needed)."""
    src_time = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                   if f.endswith(('.cu', '.cuh', '.h', 'Makefile')))
    inc = os.path.join(os.path.dirname(_HERE), 'include', 'posefit.h')
    if os.path.exists(inc):
        src_time = max(src_time, os.path.getmtime(inc))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < src_time:
        cmd = ['make', '-C', CSRC, '-B' if force else '-s', 'all']
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout, res.stderr)
        if res.returncode != 0:
            raise PoseFitError(f'building {LIB_PATH} failed:\n{res.stdout}\n{res.stderr}')
    return LIB_PATH


def _declare(lib):
    c = ctypes
    vp, i32, f64, sz = c.c_void_p, c.c_int, c.c_double, c.c_size_t
    lib.posefit_version.restype = i32
    lib.posefit_version.argtypes = []
    lib.posefit_error_string.restype = c.c_char_p
    lib.posefit_error_string.argtypes = [i32]
    lib.posefit_workspace_bytes.restype = sz
    lib.posefit_workspace_bytes.argtypes = [i32] * 5
    lib.posefit_launch_count.restype = c.c_ulonglong
    lib.posefit_launch_count.argtypes = []
    lib.posefit_debug_reload_env.restype = None
    lib.posefit_debug_reload_env.argtypes = []
    lib.posefit_forward.restype = i32
    lib.posefit_forward.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_forward_ransac.restype = i32
    lib.posefit_forward_ransac.argtypes = [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, f64, i32,
                                           vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_forward_ex.restype = i32
    lib.posefit_forward_ex.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_forward_ransac_ex.restype = i32
    lib.posefit_forward_ransac_ex.argtypes = [vp, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, f64, i32,
                                              vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_head_workspace_bytes.restype = sz
    lib.posefit_head_workspace_bytes.argtypes = [i32]
    lib.posefit_forward_head.restype = i32
    lib.posefit_forward_head.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp,
                                         vp, sz, vp]
    lib.posefit_backward_head.restype = i32
    lib.posefit_backward_head.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp,
                                          vp, vp, sz, vp]
    lib.posefit_backward.restype = i32
    lib.posefit_backward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_backward_workspace_bytes.restype = sz
    lib.posefit_backward_workspace_bytes.argtypes = [i32]
    lib.posefit_points_forward.restype = i32
    lib.posefit_points_forward.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_points_forward_ransac.restype = i32
    lib.posefit_points_forward_ransac.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, f64, f64, f64, i32,
                                                  vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_compact.restype = i32
    lib.posefit_compact.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.posefit_points_evaluate.restype = i32
    lib.posefit_points_evaluate.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.posefit_epilogue.restype = i32
    lib.posefit_epilogue.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, i32, i32, i32, vp, vp]
    lib.posefit_clip_mask.restype = i32
    lib.posefit_clip_mask.argtypes = [vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.posefit_sor_workspace_bytes.restype = sz
    lib.posefit_sor_workspace_bytes.argtypes = [i32, i32, i32]
    lib.posefit_sor_mask.restype = i32
    lib.posefit_sor_mask.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f64, i32, i32, i32, i32, vp, vp, sz, vp]
    lib.posefit_resample_noc.restype = i32
    lib.posefit_resample_noc.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.posefit_resample_noc_backward.restype = i32
    lib.posefit_resample_noc_backward.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.posefit_gather_crops.restype = i32
    lib.posefit_gather_crops.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.posefit_unpack_mask.restype = i32
    lib.posefit_unpack_mask.argtypes = [vp, c.c_longlong, vp, vp]
    lib.posefit_edge_workspace_bytes.restype = sz
    lib.posefit_edge_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.posefit_edge_features.restype = i32
    lib.posefit_edge_features.argtypes = [vp, vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, c.c_longlong,
                                          vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.posefit_transform_points.restype = i32
    lib.posefit_transform_points.argtypes = [vp, i32, vp, vp, i32, i32, vp]


def lib():
    """The loaded library (loads, and if needed builds, on first use)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    build()
                handle = ctypes.CDLL(LIB_PATH)
                missing = [s for s in SYMBOLS if not hasattr(handle, s)]
                if missing:
                    raise PoseFitError(f'{LIB_PATH} lacks symbols {missing}; rebuild with make -C {CSRC}')
                _declare(handle)
                if handle.posefit_version() != ABI_VERSION:
                    raise PoseFitError(f'ABI version mismatch: library {handle.posefit_version()}, binding {ABI_VERSION}')
                _lib = handle
    return _lib


def reload_knobs():
    """Re-read the POSEFIT_* environment knobs (the library reads them once at load)."""
    lib().posefit_debug_reload_env()


def check(code: int, what: str):
    if code != 0:
        msg = lib().posefit_error_string(code).decode()
        raise PoseFitError(f'{what} failed ({code}): {msg}')
