"""zedo_release_b200 -- B200-native implementation of ZeDO's per-pose optimisation loop.

The package holds only what the hot path needs: ``csrc/`` (hand-written sm_100a CUDA kernels
and the C ABI of ``include/zedo_b200.h``), ``engine`` (torch tensors -> C ABI) and ``lib/``
(a mirror of the reference's ``lib.*`` module paths for the functions on the path, so the
reference's drivers can run against it).  There is no CPU fallback.
"""
from . import _native  # noqa: F401  (fails loudly when libzedo_b200.so has not been built)
from .engine import (ScorePlan, aggregate_errors, axes_mask, eval_multi, grad_field, hypothesis_std,  # noqa: F401
                     ipo_fit, linspace_schedule, pck_auc, rotopt_backward, rotopt_forward, run_pose_optimisation,
                     shard_range)

__version__ = "0.1.0"
