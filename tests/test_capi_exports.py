"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports exactly the symbols include/amc3d.h declares (no compute calls without a GPU); the
product never routes through the oracle; the Tier-1/2/3 names the reference exposes exist."""
import os
import re
import subprocess

import pytest

from _util import REPO


def _declared():
    text = open(os.path.join(REPO, "include", "amc3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(amc3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from amcontrast3d_b200 import _build, _capi
    path = _build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (amc3d_[a-z0-9_]+)", out))
    declared = _declared()
    assert len(declared) >= 27
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert exported <= set(declared), f"exported but not in the header: {sorted(exported - set(declared))}"
    assert set(_capi.SIGNATURES) == set(declared)
    lib = _capi.load()
    assert lib.amc3d_version() == 100 and lib.amc3d_arch() == b"sm_100a"


def test_library_is_sm100a_only():
    from amcontrast3d_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_are_reported_not_fatal():
    from amcontrast3d_b200 import _capi
    lib = _capi.load()
    assert lib.amc3d_knnquery(10, 10, 1, 500, 0, 0, 0, 0, 0, 0, 0) == -2      # AMC3D_ELIMIT, no launch
    assert b"nsample" in lib.amc3d_last_error()
    assert lib.amc3d_furthest_point_sampling(1, 0, 4, 0, 0, 0, 0) == -1        # AMC3D_EINVAL
    with pytest.raises(_capi.Amc3dError):
        _capi.call("amc3d_stage_labels", 4, 0, 1000, 0, 0, 0, 0, 0, 0)


def test_no_cpu_fallback_in_the_product():
    import torch
    from amcontrast3d_b200 import _capi, pointops
    x = torch.rand(16, 3)
    o = torch.tensor([16], dtype=torch.int32)
    with pytest.raises(_capi.Amc3dError):
        pointops.knnquery(4, x, x, o, o)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "amcontrast3d_b200")
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "ops_oracle" not in src and "loss_oracle" not in src, fn


def test_reference_facing_names_exist():
    from amcontrast3d_b200 import layers, loss, pointnet2_batch_cuda, pointops, pointops_cuda
    from amcontrast3d_b200.AMContrast3D import AmbiguityHead, ContrastHead, RefinementMethod, posmask_searching
    from amcontrast3d_b200.AMContrast3D.AEF.ambiguity import ambiguity_function
    from amcontrast3d_b200.AMContrast3D.AEF.utils import fetch_pxo, get_subscene_label_CBL
    for n in ("ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
              "gather_points_grad_wrapper", "furthest_point_sampling_wrapper", "three_nn_wrapper",
              "three_interpolate_wrapper", "three_interpolate_grad_wrapper"):     # pointnet2_api.cpp:10-24
        assert callable(getattr(pointnet2_batch_cuda, n))
    for n in ("knnquery_cuda", "ballquery_cuda", "furthestsampling_cuda", "grouping_forward_cuda",
              "grouping_backward_cuda", "interpolation_forward_cuda", "interpolation_backward_cuda",
              "subtraction_forward_cuda", "subtraction_backward_cuda", "aggregation_forward_cuda",
              "aggregation_backward_cuda"):                                        # pointops_api.cpp:13-25
        assert callable(getattr(pointops_cuda, n))
    for n in ("furthest_point_sample", "gather_operation", "ball_query", "grouping_operation", "three_nn",
              "three_interpolate", "three_interpolation", "create_grouper", "QueryAndGroup", "KNNGroup", "GroupAll"):
        assert hasattr(layers, n)
    assert callable(pointops.knnquery)
    assert {"CrossEntropyAce", "CrossEntropyAcePre"} <= set(loss.LOSS)
    # loss / ambiguity heads must stay parameter- and buffer-free (checkpoint compatibility)
    assert len(ContrastHead().state_dict()) == 0 and len(AmbiguityHead().state_dict()) == 0
    assert ContrastHead().stages == [('up', 0), ('up', 1), ('up', 2), ('up', 3)]


def test_division_free_index_arithmetic_of_the_fused_forward():
    """FastDiv (csrc/fused_sa.cu): multiply-high magic for n / d, n < 2^31 — every divisor the kernel can meet
    (queries per scene, output-channel slices) against integer division, on the host copy of the same code"""
    import random
    from amcontrast3d_b200 import _capi
    f = _capi.load().amc3d_debug_fastdiv
    rng = random.Random(7)
    divisors = list(range(1, 1200)) + [2 ** k + d for k in range(1, 25) for d in (-1, 0, 1)] + \
        [rng.randrange(1, 1 << 25) for _ in range(1500)] + [6000, 24000, 93, 375, 1500, 16000, 64000]
    for d in divisors:
        if d < 1:
            continue
        for n in (0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 7 * d + 3, (1 << 31) - 1, (1 << 31) - d, (1 << 25) - 1,
                  rng.randrange(0, 1 << 31), rng.randrange(0, 1 << 25)):
            if 0 <= n < (1 << 31):
                assert f(n, d) == n // d, (n, d)
