"""oracle/ops_oracle.c against outputs of the REFERENCE's own CUDA kernels run on a B200
(tests/golden/ref_kernels_golden.npz, produced by tests/golden/make_ref_kernel_golden.py from
oracle/_ref).  Integer / copy outputs must be identical, including the literal heap and
tree-reduction behaviour under exact ties."""
import os
import sys

import numpy as np
import pytest

from _util import REPO, rel_err
from oracle import ops_oracle as oo

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
PATH = os.path.join(REPO, "tests", "golden", "ref_kernels_golden.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="ref_kernels_golden.npz not generated yet")


@pytest.fixture(scope="module")
def data():
    from make_ref_kernel_golden import golden_inputs
    return golden_inputs(), np.load(PATH)


def test_fps(data):
    inp, g = data
    for m in (750, 100):
        idx, temp = oo.fps(inp["xyz"], m)
        assert np.array_equal(idx, g[f"fps/{m}/idx"]) and np.array_equal(temp, g[f"fps/{m}/temp"])
    idx, temp = oo.fps(inp["lattice"], 400)
    assert np.array_equal(idx, g["fps/lattice/idx"]) and np.array_equal(temp, g["fps/lattice/temp"])


def test_ball_query_three_nn_interpolate(data):
    inp, g = data
    xyz = inp["xyz"]
    q = np.ascontiguousarray(xyz[:, :750])
    for r, ns in ((0.1, 32), (0.2, 16)):
        assert np.array_equal(oo.ball_query(r, ns, xyz, q), g[f"ball_query/{r}_{ns}"])
    d2, i3 = oo.three_nn(xyz, q)
    assert np.array_equal(i3, g["three_nn/idx"]) and np.array_equal(d2, g["three_nn/dist2"])
    out = oo.three_interpolate(np.ascontiguousarray(inp["feats"][:, :, :750]), i3, inp["w"])
    assert np.array_equal(out, g["three_interpolate/out"])


def test_grouping(data):
    inp, g = data
    assert np.array_equal(oo.group_points(inp["feats"], inp["gidx"]), g["group_points/out"])
    assert rel_err(oo.group_points_grad(inp["go"], inp["gidx"], 3000), g["group_points_grad/out"]) <= 1e-6


def test_knn(data):
    inp, g = data
    flat = np.ascontiguousarray(inp["xyz"].reshape(-1, 3))
    o1 = np.array([6000], dtype=np.int32)
    for k in (4, 16, 24, 64):
        i, d = oo.knnquery(k, flat, flat, o1, o1)
        assert np.array_equal(i, g[f"knn/{k}/idx"]) and np.array_equal(d, g[f"knn/{k}/dist2"])
    so = np.concatenate([inp["segs"], [6000]]).astype(np.int32)
    i, d = oo.knnquery(12, flat, flat, so, so)
    assert np.array_equal(i, g["knn/segs/idx"]) and np.array_equal(d, g["knn/segs/dist2"])
    lat = np.ascontiguousarray(inp["lattice"].reshape(-1, 3))
    ol = np.array([3000], dtype=np.int32)
    i, d = oo.knnquery(8, lat, lat, ol, ol)
    assert np.array_equal(d, g["knn/lattice/dist2"])
    assert np.array_equal(i, g["knn/lattice/idx"])          # literal heap: same order under ties
