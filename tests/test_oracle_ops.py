"""CPU checks of the operator oracle itself (oracle/ops_oracle.c): internal consistency and the
semantics documented in SURVEY.md App. A, at sizes that finish in seconds."""
import numpy as np

from amcontrast3d_b200 import scenes
from oracle import ops_oracle as oo


def test_opt_n_threads_table():
    # cuda_utils.h:10-14 (values verified against the reference in SURVEY.md App. B)
    for n, bs in ((24000, 1024), (6000, 1024), (1500, 1024), (1000, 512), (375, 256), (250, 128), (93, 64),
                  (2048, 1024), (1024, 1024), (512, 512), (8, 8), (1, 1)):
        assert oo.opt_n_threads(n) == bs


def test_knn_heap_equals_lexicographic_on_tie_free_input():
    xyz, _ = scenes.surface_scene(3000, seed=1)
    o = np.array([3000], dtype=np.int32)
    i1, d1 = oo.knnquery(16, xyz, None, o, o)
    i2, d2 = oo.knnquery(16, xyz, None, o, o, lex=True)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert (i1[:, 0] == np.arange(3000)).all() and (d1[:, 0] == 0).all()
    # brute force in float64 agrees on membership
    d = ((xyz[:200, None, :].astype(np.float64) - xyz[None].astype(np.float64)) ** 2).sum(-1)
    assert np.array_equal(np.sort(np.argsort(d, axis=1)[:, :16], axis=1), np.sort(i1[:200], axis=1))


def test_knn_short_segment_padding():
    xyz = np.random.default_rng(0).random((10, 3), dtype=np.float32)
    o = np.array([4, 10], dtype=np.int32)
    idx, d2 = oo.knnquery(6, xyz, None, o, o)
    assert (idx[:4, 4:] == 0).all() and (d2[:4, 4:] == np.float32(1e10)).all()     # (start, 1e10) tail
    assert (idx[4:] >= 4).all()


def test_fps_properties():
    xyz, _ = scenes.batch_of_scenes(2, 1500, "surface")
    idx, temp = oo.fps(xyz, 375)
    assert (idx[:, 0] == 0).all()
    for b in range(2):
        assert len(set(idx[b].tolist())) == 375
        # the second pick is the farthest point from point 0
        d = ((xyz[b] - xyz[b, 0]) ** 2).sum(1)
        assert d[idx[b, 1]] == d.max()
    assert (temp[np.arange(2)[:, None], idx[:, :-1]] == 0).all()


def test_ball_query_semantics():
    xyz, _ = scenes.batch_of_scenes(1, 2000, "surface")
    q = xyz[:, :300]
    idx = oo.ball_query(0.1, 16, xyz, q)
    d = ((q[0, :, None, :] - xyz[0, None]) ** 2).sum(-1)
    for j in range(0, 300, 37):
        hits = np.nonzero(d[j] < np.float32(0.1) ** 2)[0][:16]
        exp = np.full(16, hits[0])
        exp[:len(hits)] = hits
        assert np.array_equal(idx[0, j], exp)
    far = oo.ball_query(0.01, 8, xyz, q + 50.0)
    assert (far == 0).all()                       # no hit: row left at the caller's zeros


def test_group_and_interpolate_adjointness():
    rng = np.random.default_rng(0)
    f = rng.standard_normal((2, 5, 100)).astype(np.float32)
    idx = rng.integers(0, 100, size=(2, 30, 4)).astype(np.int32)
    out = oo.group_points(f, idx)
    go = rng.standard_normal(out.shape).astype(np.float32)
    gf = oo.group_points_grad(go, idx, 100)
    assert np.isclose((out.astype(np.float64) * go).sum(), (f.astype(np.float64) * gf).sum(), rtol=1e-5)
    w = rng.random((2, 30, 3)).astype(np.float32)
    i3 = rng.integers(0, 100, size=(2, 30, 3)).astype(np.int32)
    o3 = oo.three_interpolate(f, i3, w)
    g3 = rng.standard_normal(o3.shape).astype(np.float32)
    gf3 = oo.three_interpolate_grad(g3, i3, w, 100)
    assert np.isclose((o3.astype(np.float64) * g3).sum(), (f.astype(np.float64) * gf3).sum(), rtol=1e-5)
