"""B200-native per-object 7-DoF pose solver (drop-in for PoseEst/ of
DomiSchmauser/3D_MOT_Differentiable_Pose_Estimation).

The package name starts with a digit, so import it with
`importlib.import_module('3d_mot_differentiable_pose_estimation_b200')`.

Layout: `csrc/` holds the sm_100a kernels and the C ABI (include/posefit.h); `function.py` the
torch.autograd.Function over it; `pose_utils.py` / `pose_estimation.py` mirror the reference's
own modules (same function names and signatures); `synth.py` makes MOTFront-shaped inputs;
`shard.py` partitions objects by sequence across GPUs.
"""
from . import _lib  # noqa: F401
from .function import (PoseFit, PoseFitFull, PoseFitRaw, pose_fit, pose_fit_raw, points_fit_raw,  # noqa: F401
                       pose_fit_backward_raw, default_kinv, pose_epilogue, PoseEpilogue, clip_mask_to_box, statistical_outlier_mask,
                       STATUS_OK, STATUS_EMPTY, STATUS_LOW_INLIER_RATIO, STATUS_NAN, REF_COMPAT, SAMPLES_ARE_BITS,
                       device_sample_bits, PoseFitHead, pose_fit_head, ransac_iterations)
from . import synth, shard  # noqa: F401
from .frontend import (gather_crops, resample_noc, ResampleNoc, Crops, run_pose_batched, BatchedPoses,  # noqa: F401
                       pack_mask, unpack_mask)
from . import graph_dataset  # noqa: F401  (drop-in for Tracking/datasets/graph_dataset.py's edge construction)
from . import pose_utils, pose_estimation  # noqa: F401  (drop-ins for PoseEst/pose_utils.py, pose_estimation.py)

__all__ = ['PoseFit', 'PoseFitRaw', 'ransac_iterations', 'pose_fit', 'pose_fit_raw', 'points_fit_raw', 'pose_fit_backward_raw',
           'default_kinv', 'gather_crops', 'pack_mask', 'unpack_mask', 'resample_noc', 'run_pose_batched', 'BatchedPoses', 'pose_epilogue', 'PoseEpilogue', 'clip_mask_to_box', 'statistical_outlier_mask', 'synth', 'STATUS_OK', 'STATUS_EMPTY', 'STATUS_LOW_INLIER_RATIO', 'STATUS_NAN',
           'REF_COMPAT', 'SAMPLES_ARE_BITS', 'device_sample_bits', 'PoseFitHead', 'pose_fit_head']
