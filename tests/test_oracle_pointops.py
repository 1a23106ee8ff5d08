"""CPU checks of the oracle_pop_* restatements (the remaining pointops kernels) against independent numpy
formulations, and of their segment handling against the batched restatements that are already pinned to the
reference's kernels."""
import numpy as np

from _util import rel_err
from oracle import ops_oracle as oo


def _case(seed=0, n=300, ns=6, c=10, w_c=5):
    rng = np.random.default_rng(seed)
    return dict(a=rng.standard_normal((n, c)).astype(np.float32), b=rng.standard_normal((n, c)).astype(np.float32),
                idx=rng.integers(0, n, size=(n, ns)).astype(np.int32),
                pos=rng.standard_normal((n, ns, c)).astype(np.float32),
                w=rng.standard_normal((n, ns, w_c)).astype(np.float32),
                g=rng.standard_normal((n, c)).astype(np.float32),
                g3=rng.standard_normal((n, ns, c)).astype(np.float32))


def test_furthestsampling_segments_match_batched_restatement():
    rng = np.random.default_rng(1)
    xyz = rng.random((1100, 3)).astype(np.float32)
    # segments of 300 and 800 points; n_max = 800 -> 512 threads, the block size the batched kernel uses for n = 800
    off, noff = np.array([300, 1100], np.int32), np.array([40, 240], np.int32)
    idx = oo.pop_furthestsampling(xyz, off, noff)
    ref1, _ = oo.fps(xyz[None, 300:], 200)
    assert np.array_equal(idx[40:] - 300, ref1[0])
    assert idx[0] == 0 and idx[40] == 300 and (idx[:40] < 300).all() and len(set(idx[:40].tolist())) == 40
    # equal-sized segments: every segment equals the batched result
    off, noff = np.array([550, 1100], np.int32), np.array([100, 200], np.int32)
    idx = oo.pop_furthestsampling(xyz, off, noff)
    ref, _ = oo.fps(xyz.reshape(2, 550, 3), 100)
    assert np.array_equal(idx.reshape(2, 100) - np.array([[0], [550]]), ref)


def test_ballquery_against_numpy_and_batched_restatement():
    rng = np.random.default_rng(2)
    xyz = rng.random((900, 3)).astype(np.float32)
    q = rng.random((120, 3)).astype(np.float32)
    off, qoff = np.array([400, 900], np.int32), np.array([50, 120], np.int32)
    got = oo.pop_ballquery(0.15, 8, xyz, q, off, qoff)
    for s, (a0, a1, q0, q1) in enumerate(((0, 400, 0, 50), (400, 900, 50, 120))):
        ref = oo.ball_query(0.15, 8, xyz[None, a0:a1], q[None, q0:q1])[0]
        hit = (got[q0:q1] != 0).any(1) | (ref != 0).any(1)
        assert np.array_equal(got[q0:q1][hit] - a0, ref[hit])
    # a radius nobody reaches leaves the zero fill
    assert not oo.pop_ballquery(1e-6, 4, xyz, q, off, qoff).any()


def test_interpolation_subtraction_aggregation_formulas():
    d = _case()
    a, b, idx, pos, w, g, g3 = (d[k] for k in ("a", "b", "idx", "pos", "w", "g", "g3"))
    n, ns = idx.shape
    c, w_c = a.shape[1], w.shape[2]
    li = idx.astype(np.int64)
    assert np.array_equal(oo.pop_subtraction_fwd(a, b, idx), a[:, None, :] - b[li])
    g1, g2 = oo.pop_subtraction_bwd(idx, g3)
    e2 = np.zeros((n, c), np.float64)
    np.add.at(e2, li.reshape(-1), -g3.reshape(-1, c).astype(np.float64))
    assert rel_err(g1, g3.astype(np.float64).sum(1)) < 1e-6 and rel_err(g2, e2) < 1e-6
    wt = np.tile(w, (1, 1, c // w_c)).astype(np.float64)
    exp = ((a[li].astype(np.float64) + pos) * wt).sum(1)
    assert rel_err(oo.pop_aggregation_fwd(a, pos, w, idx), exp) < 1e-6
    gi, gp, gw = oo.pop_aggregation_bwd(a, pos, w, idx, g)
    assert np.array_equal(gp, g[:, None, :] * np.tile(w, (1, 1, c // w_c)))
    ei = np.zeros((n, c), np.float64)
    np.add.at(ei, li.reshape(-1), (g[:, None, :].astype(np.float64) * wt).reshape(-1, c))
    assert rel_err(gi, ei) < 1e-6
    ew = (g[:, None, :].astype(np.float64) * (a[li] + pos)).reshape(n, ns, c // w_c, w_c).sum(2)
    assert rel_err(gw, ew) < 1e-6
    k = 3
    iw = np.abs(w[:, :k, 0]).copy()
    out = oo.pop_interpolation_fwd(a, idx[:, :k].copy(), iw)
    assert rel_err(out, (a[li[:, :k]].astype(np.float64) * iw[:, :, None]).sum(1)) < 1e-6
    # accumulation into a caller-provided output
    assert rel_err(oo.pop_interpolation_fwd(a, idx[:, :k].copy(), iw, output=b), out.astype(np.float64) + b) < 1e-6
    eb = np.zeros((n, c), np.float64)
    np.add.at(eb, li[:, :k].reshape(-1), (g[:, None, :].astype(np.float64) * iw[:, :, None]).reshape(-1, c))
    assert rel_err(oo.pop_interpolation_bwd(g, idx[:, :k].copy(), iw, n), eb) < 1e-6
