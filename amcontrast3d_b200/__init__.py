"""amcontrast3d_b200 — B200-native (sm_100a) point-grouping operators and adaptive-margin
contrastive loss behind AMContrast3D's own Python operator / loss-module API.

Layout (only what the hot path needs):
  csrc/, lib/                 CUDA kernels + the C-ABI of include/amc3d.h, built in-tree by _build.py
  _capi.py                    ctypes door to the library (no CPU fallback)
  pointnet2_batch_cuda.py,    Tier 1: drop-ins for the reference's two extension modules
  pointops_cuda.py
  layers/, pointops.py        Tier 2: furthest_point_sample, ball_query, grouping_operation,
                              three_nn / three_interpolate, knnquery with the reference signatures
  AMContrast3D/, loss.py      Tier 3: ContrastHead, AmbiguityHead, RefinementMethod, criteria
  compat.py                   install the above under the reference's import paths
  scenes.py, replay.py, dist.py   synthetic inputs, the timed path replay, data-parallel harness
"""
__version__ = "0.1.0"

from . import _capi  # noqa: F401  (does not load the library until first use)


def load_library():
    """Load (building if necessary) libamc3d_sm100a.so; raises if that is impossible."""
    return _capi.load()
