"""compat.install puts the sm_100a library behind the reference's extension-module names (Tier 1)."""
import importlib
import os
import sys

import pytest


def test_install_tier1_registers_extension_modules():
    from amcontrast3d_b200 import compat, pointnet2_batch_cuda, pointops_cuda
    saved = {k: sys.modules.get(k) for k in ("pointnet2_batch_cuda", "pointops_cuda")}
    try:
        done = compat.install(tier=1)
        assert "pointnet2_batch_cuda" in done and "pointops_cuda" in done
        assert importlib.import_module("pointops_cuda") is pointops_cuda
        assert importlib.import_module("pointnet2_batch_cuda") is pointnet2_batch_cuda
        # the names the reference's pybind tables export on the hot path (pointnet2_api.cpp:10-24, pointops_api.cpp:14)
        for name in ("ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
                     "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "three_nn_wrapper",
                     "three_interpolate_wrapper", "three_interpolate_grad_wrapper"):
            assert callable(getattr(pointnet2_batch_cuda, name))
        # every export of pointops_api.cpp:13-25
        for name in ("furthestsampling_cuda", "knnquery_cuda", "ballquery_cuda", "grouping_forward_cuda",
                     "grouping_backward_cuda", "interpolation_forward_cuda", "interpolation_backward_cuda",
                     "subtraction_forward_cuda", "subtraction_backward_cuda", "aggregation_forward_cuda",
                     "aggregation_backward_cuda"):
            assert callable(getattr(pointops_cuda, name))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.isdir("/root/reference/openpoints"), reason="needs the reference checkout")
def test_reference_wrappers_bind_to_the_library():
    """The reference's own pointops.py imports `pointops_cuda` by name: after install() that is ours."""
    from amcontrast3d_b200 import compat, pointops_cuda
    saved_path = list(sys.path)
    saved = {k: sys.modules.get(k) for k in ("pointnet2_batch_cuda", "pointops_cuda")}
    try:
        compat.install(tier=1)
        sys.path.insert(0, "/root/reference/openpoints/cpp/pointops/functions")
        ref_pointops = importlib.import_module("pointops")
        assert ref_pointops.pointops_cuda is pointops_cuda
        assert hasattr(ref_pointops, "knnquery")
    finally:
        sys.path[:] = saved_path
        sys.modules.pop("pointops", None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
