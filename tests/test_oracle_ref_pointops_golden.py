"""The oracle_pop_* restatements (oracle/ops_oracle.c) against outputs of the REFERENCE's own CUDA kernels run
on a B200 (tests/golden/ref_pointops_golden.npz, produced by tests/golden/make_ref_pointops_golden.py from
oracle/_ref/libref_pointops.so).  Index, gather and single-FMA-chain outputs must be identical; atomically
accumulated gradients agree to rounding."""
import os
import sys

import numpy as np
import pytest

from _util import REPO, rel_err
from oracle import ops_oracle as oo

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
PATH = os.path.join(REPO, "tests", "golden", "ref_pointops_golden.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="ref_pointops_golden.npz not generated yet")


@pytest.fixture(scope="module")
def data():
    from make_ref_pointops_golden import golden_inputs
    return golden_inputs(), np.load(PATH)


def test_furthestsampling(data):
    inp, g = data
    assert np.array_equal(oo.pop_furthestsampling(inp["xyz"], inp["offset"], inp["new_offset"]), g["fps/idx"])
    assert np.array_equal(oo.pop_furthestsampling(inp["lattice"], inp["lat_offset"], inp["lat_new_offset"]),
                          g["fps/lattice/idx"])


def test_ballquery(data):
    inp, g = data
    xyz, off, noff = inp["xyz"], inp["offset"], inp["new_offset"]
    q = np.ascontiguousarray(xyz[g["fps/idx"].astype(np.int64)])
    for r, ns in ((0.1, 16), (0.25, 8), (0.02, 4)):
        assert np.array_equal(oo.pop_ballquery(r, ns, xyz, q, off, noff), g[f"ballquery/{r}_{ns}"])
    lat, loff = inp["lattice"], inp["lat_offset"]
    assert np.array_equal(oo.pop_ballquery(0.5, 12, lat, lat, loff, loff), g["ballquery/lattice"])


def test_interpolation_subtraction_aggregation(data):
    inp, g = data
    assert np.array_equal(oo.pop_interpolation_fwd(inp["src"], inp["iidx"], inp["iw"]), g["interpolation/fwd"])
    assert rel_err(oo.pop_interpolation_bwd(inp["g_nc"], inp["iidx"], inp["iw"], inp["src"].shape[0]),
                   g["interpolation/bwd"]) < 1e-6
    assert np.array_equal(oo.pop_subtraction_fwd(inp["feat"], inp["feat2"], inp["nidx"]), g["subtraction/fwd"])
    g1, g2 = oo.pop_subtraction_bwd(inp["nidx"], inp["g_nsc"])
    assert rel_err(g1, g["subtraction/g1"]) < 1e-6 and rel_err(g2, g["subtraction/g2"]) < 1e-6
    assert np.array_equal(oo.pop_aggregation_fwd(inp["feat"], inp["pos"], inp["wgt"], inp["nidx"]),
                          g["aggregation/fwd"])
    gi, gp, gw = oo.pop_aggregation_bwd(inp["feat"], inp["pos"], inp["wgt"], inp["nidx"], inp["g_nc"])
    assert np.array_equal(gp, g["aggregation/gp"])
    assert rel_err(gi, g["aggregation/gi"]) < 1e-6 and rel_err(gw, g["aggregation/gw"]) < 1e-6
