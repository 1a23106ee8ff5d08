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


def test_install_tier4_routes_module_forwards_through_the_fused_operator():
    """tier 4 wraps LocalAggregation.forward / SetAbstraction.forward of the backbone modules; configurations the
    operator does not cover (here: CPU tensors) fall through to the original forward"""
    import types
    import torch
    from amcontrast3d_b200 import compat
    name = "openpoints.models.backbone.pointnext_AA"
    saved = {k: sys.modules.get(k) for k in ("pointnet2_batch_cuda", "pointops_cuda", name)}

    class LocalAggregation(torch.nn.Module):
        def forward(self, pf):
            return "reference composition"

    class SetAbstraction(LocalAggregation):
        pass

    fake = types.ModuleType(name)
    fake.LocalAggregation, fake.SetAbstraction = LocalAggregation, SetAbstraction
    try:
        sys.modules[name] = fake
        done = compat.install(tier=4)
        assert f"{name}.LocalAggregation.forward" in done and f"{name}.SetAbstraction.forward" in done
        assert LocalAggregation._amc3d_fused and SetAbstraction._amc3d_fused
        m = LocalAggregation()
        m.grouper = types.SimpleNamespace(radius=0.1, nsample=32, normalize_dp=True)
        m.feature_type, m.reduction = "dp_fj", "max"
        m.convs = torch.nn.Sequential(torch.nn.Sequential(torch.nn.Conv2d(35, 32, 1, bias=False), torch.nn.BatchNorm2d(32),
                                                          torch.nn.ReLU()))
        assert m((torch.zeros(1, 8, 3), torch.zeros(1, 32, 8))) == "reference composition"     # CPU tensors: not fusable
        compat.install(tier=4)                                                                # idempotent
        assert LocalAggregation.forward is not LocalAggregation._amc3d_orig_forward
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
