"""kNN tie note: the reference heap (knnquery_cuda_kernel.cu:21-48) is order- and membership-unstable
under exactly equal distances; the sm_100a kernel returns the (d2, index)-lexicographic top-k
("ties broken by lowest index", BASELINE.json north_star).  The two agree on every tie-free row.
Random float32 ties do occur at scale (~k^2/2^24 per row), so: squared distances are compared
bit-exactly against the literal-heap oracle on ALL rows (the sorted multiset is tie-independent),
indices bit-exactly against the lexicographic oracle on all rows, and against the literal-heap
oracle on the rows whose k+1 nearest distances are distinct.

GPU parity of the point-grouping operators against the CPU restatement of the reference
kernels (oracle/ops_oracle.c), called through the reference-facing Python operators, which in
turn go through the C-ABI of include/amc3d.h.  Bar: bit-exact indices / distances / gathers;
scatter-add gradients within 1e-5 relative (the reference's own atomics are order-dependent)."""
import os

import numpy as np
import pytest
import torch

from _util import REPO, rel_err
from amcontrast3d_b200 import scenes
from oracle import ops_oracle as oo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ------------------------------------------------------------------ kNN
def _check_knn(idx, dist, k, xyz, q, o, qo, sqrt):
    hi, hd2 = oo.knnquery(k, xyz, q, o, qo)                 # literal reference heap
    li, ld2 = oo.knnquery(k, xyz, q, o, qo, lex=True)       # (d2, index) lexicographic
    exp = np.sqrt(hd2) if sqrt else hd2
    assert np.array_equal(dist, exp)
    assert np.array_equal(idx, li)
    if k < 128:
        _, d2p = oo.knnquery(k + 1, xyz, q, o, qo)
        tie_free = (np.diff(d2p, axis=1) > 0).all(1) | (d2p[:, -1] >= 1e10)
        tie_free &= (np.diff(hd2, axis=1) > 0).all(1) | (hd2[:, -1] >= 1e10)
        assert tie_free.mean() > 0.98
        assert np.array_equal(idx[tie_free], hi[tie_free])


@pytest.mark.parametrize("k", [1, 3, 4, 8, 12, 16, 24, 32, 33, 64, 100])
def test_knn_single_segment(k):
    from amcontrast3d_b200 import pointops
    xyz, _ = scenes.surface_scene(6000, seed=3)
    o = np.array([6000], dtype=np.int32)
    idx, dist = pointops.knnquery(k, _t(xyz), None, _t(o), _t(o))
    assert idx.dtype == torch.int32 and dist.dtype == torch.float32
    _check_knn(idx.cpu().numpy(), dist.cpu().numpy(), k, xyz, xyz, o, o, sqrt=True)


def test_knn_raw_dist2_bit_exact_and_nsample_tensor():
    from amcontrast3d_b200 import _amloss, pointops
    xyz, _ = scenes.volume_scene(5000, seed=5)
    o = np.array([5000], dtype=np.int32)
    ri, rd2 = oo.knnquery(16, xyz, None, o, o, lex=True)
    idx, d2 = _amloss.knn_raw(16, _t(xyz), None, _t(o), _t(o))
    _check_knn(idx.cpu().numpy(), d2.cpu().numpy(), 16, xyz, xyz, o, o, sqrt=False)
    # nsample as a 0-dim tensor (AEF/utils.py:29 passes torch.prod(...))
    idx2, _ = pointops.knnquery(torch.prod(torch.tensor([4, 4])), _t(xyz), _t(xyz), _t(o), _t(o))
    assert np.array_equal(idx2.cpu().numpy(), ri)


def test_knn_ragged_segments_and_cross_sets():
    from amcontrast3d_b200 import pointops
    rng = np.random.default_rng(0)
    sizes = [700, 5, 1300, 259]            # one segment shorter than k
    qsizes = [300, 40, 0, 513]             # an empty query segment
    xyz = rng.random((sum(sizes), 3), dtype=np.float32) * 3
    q = rng.random((sum(qsizes), 3), dtype=np.float32) * 3
    o = np.cumsum(sizes).astype(np.int32)
    qo = np.cumsum(qsizes).astype(np.int32)
    for k in (8, 16, 40):
        ri, rd2 = oo.knnquery(k, xyz, q, o, qo, lex=True)
        idx, dist = pointops.knnquery(k, _t(xyz), _t(q), _t(o), _t(qo))
        assert np.array_equal(idx.cpu().numpy(), ri)
        assert np.array_equal(dist.cpu().numpy(), np.sqrt(rd2))


def test_knn_label_vote_shape():
    """stage-0 support, sub-sampled queries, kr = 4/16/64 as in AEF/utils.py:29-36"""
    from amcontrast3d_b200 import pointops
    xyz, _ = scenes.surface_scene(8192, seed=9)
    for kr, m in ((4, 2048), (16, 512), (64, 128)):
        q = np.ascontiguousarray(xyz[:m])
        o, qo = np.array([8192], dtype=np.int32), np.array([m], dtype=np.int32)
        idx, dist = pointops.knnquery(kr, _t(xyz), _t(q), _t(o), _t(qo))
        _check_knn(idx.cpu().numpy(), dist.cpu().numpy(), kr, xyz, q, o, qo, sqrt=True)


def test_knn_empty_and_errors():
    from amcontrast3d_b200 import _capi, pointops
    xyz = _t(np.random.default_rng(1).random((100, 3), dtype=np.float32))
    o = _t(np.array([100], dtype=np.int32))
    idx, dist = pointops.knnquery(4, xyz, xyz[:0].contiguous(), o, _t(np.array([0], dtype=np.int32)))
    assert idx.shape == (0, 4)
    with pytest.raises(_capi.Amc3dError):
        pointops.knnquery(200, xyz, xyz, o, o)        # beyond the documented nsample limit
    with pytest.raises(_capi.Amc3dError):
        pointops.knnquery(4, xyz.cpu(), None, o.cpu(), o.cpu())   # no CPU path


def test_knn_full_size_properties():
    """BASELINE config 2 size (8 x 24 000 points flattened into one segment): sortedness, self
    match, and a sample of queries against the oracle."""
    from amcontrast3d_b200 import _amloss
    xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
    flat = np.ascontiguousarray(xyz.reshape(-1, 3))
    n = flat.shape[0]
    o = np.array([n], dtype=np.int32)
    idx, d2 = _amloss.knn_raw(16, _t(flat), None, _t(o), _t(o))
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    assert (np.diff(d2, axis=1) >= 0).all()
    assert (d2[:, 0] == 0).all()
    sample = np.random.default_rng(0).choice(n, 1024, replace=False)
    _check_knn(idx[sample], d2[sample], 16, flat, np.ascontiguousarray(flat[sample]), o,
               np.array([1024], dtype=np.int32), sqrt=False)


def test_knn_whole_room_size():
    """eval-time shape (metrics.py:160-184 runs the kNN on a whole room as ONE segment): 400 000 points, checked
    through sortedness, the self match, and 512 sampled queries against the oracle's brute force"""
    from amcontrast3d_b200 import _amloss
    xyz, _ = scenes.volume_scene(400000, seed=21)
    n = xyz.shape[0]
    o = np.array([n], dtype=np.int32)
    idx, d2 = _amloss.knn_raw(16, _t(xyz), None, _t(o), _t(o))
    idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
    assert (np.diff(d2, axis=1) >= 0).all()
    assert (d2[:, 0] == 0).all() and (idx[:, 0] == np.arange(n)).mean() > 0.999
    sample = np.random.default_rng(1).choice(n, 512, replace=False)
    _check_knn(idx[sample], d2[sample], 16, xyz, np.ascontiguousarray(xyz[sample]), o,
               np.array([512], dtype=np.int32), sqrt=False)


# ------------------------------------------------------------------ FPS
@pytest.mark.parametrize("n,m,b", [(93, 23, 3), (375, 93, 2), (512, 128, 1), (1500, 375, 3), (2048, 512, 2),
                                   (2049, 300, 2), (6000, 1500, 2), (8192, 600, 1), (24000, 6000, 2),
                                   (24576, 400, 1)])
def test_fps_matches_reference_sequence(n, m, b):
    from amcontrast3d_b200.layers import furthest_point_sample
    xyz, _ = scenes.batch_of_scenes(b, n, "surface", first_scene=n % 17)
    ridx, _ = oo.fps(xyz, m)
    idx = furthest_point_sample(_t(xyz), m)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (b, m)
    assert np.array_equal(idx.cpu().numpy(), ridx)


@pytest.mark.parametrize("n,m", [(64000, 300), (70000, 64), (100000, 200), (150000, 300), (200000, 200), (230000, 40)])
def test_fps_large_scenes(n, m):
    """ScanNet-sized scenes: 16-CTA cluster path and the global-memory fallback"""
    from amcontrast3d_b200.layers import furthest_point_sample
    xyz, _ = scenes.batch_of_scenes(2, n, "volume", first_scene=2)
    ridx, _ = oo.fps(xyz, m)
    idx = furthest_point_sample(_t(xyz), m)
    assert np.array_equal(idx.cpu().numpy(), ridx)


@pytest.mark.parametrize("n,m,kind,culled", [(64000, 16000, "surface", False), (50000, 3000, "volume", True),
                                             (120000, 1500, "surface", True), (230000, 800, "surface", False)])
def test_fps_long_sequences_and_culled_kernel(n, m, kind, culled, monkeypatch):
    """long pick sequences (config 3's 64 000 -> 16 000 in full) and the running distances left in `temp`, bit for
    bit; `culled` forces the spatially culled kernel (knn_grid.cu fps_culled_kernel, the default beyond 212 992
    points) onto smaller scenes as well"""
    import subprocess, sys, json
    if culled:
        # the crossover is read once per process: run this case in a child with the override
        code = ("import sys, json, numpy as np, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
                "from amcontrast3d_b200 import scenes, pointnet2_batch_cuda as ext\n"
                "from oracle import ops_oracle as oo\n"
                "xyz, _ = scenes.batch_of_scenes(2, %d, %r, first_scene=5)\n"
                "ridx, rtemp = oo.fps(xyz, %d)\n"
                "x = torch.from_numpy(xyz).cuda(); temp = torch.full((2, %d), 1e10, device='cuda')\n"
                "out = torch.zeros((2, %d), dtype=torch.int32, device='cuda')\n"
                "ext.furthest_point_sampling_wrapper(2, %d, %d, x, temp, out)\n"
                "print(json.dumps([bool(np.array_equal(out.cpu().numpy(), ridx)), bool(np.array_equal(temp.cpu().numpy(), rtemp))]))\n"
                % (REPO, os.path.join(REPO, "tests"), n, kind, m, n, m, n, m))
        env = dict(os.environ, AMC3D_FPS_CULLED_MIN="1000")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert json.loads(r.stdout.strip().splitlines()[-1]) == [True, True]
        return
    from amcontrast3d_b200 import pointnet2_batch_cuda as ext
    xyz, _ = scenes.batch_of_scenes(2, n, kind, first_scene=5)
    ridx, rtemp = oo.fps(xyz, m)
    x = _t(xyz)
    temp = torch.full((2, n), 1e10, device=DEV)
    out = torch.zeros((2, m), dtype=torch.int32, device=DEV)
    ext.furthest_point_sampling_wrapper(2, n, m, x, temp, out)
    assert np.array_equal(out.cpu().numpy(), ridx)
    assert np.array_equal(temp.cpu().numpy(), rtemp)


@pytest.mark.parametrize("n,m", [(300, 120), (1500, 500), (5000, 700), (24000, 800)])
def test_fps_exact_tie_order(n, m):
    """Lattice points produce many exactly tied maxima; the winner must follow the reference's
    shared-memory tree (bit-reversed thread order), which the oracle emulates literally."""
    from amcontrast3d_b200.layers import furthest_point_sample
    rng = np.random.default_rng(n)
    xyz = rng.integers(0, 12, size=(2, n, 3)).astype(np.float32) * 0.25
    ridx, rtemp = oo.fps(xyz, m)
    from amcontrast3d_b200 import pointnet2_batch_cuda as ext
    x = _t(xyz)
    temp = torch.full((2, n), 1e10, device=DEV)
    out = torch.zeros((2, m), dtype=torch.int32, device=DEV)
    ext.furthest_point_sampling_wrapper(2, n, m, x, temp, out)
    assert np.array_equal(out.cpu().numpy(), ridx)
    assert np.array_equal(temp.cpu().numpy(), rtemp)        # temp is left holding the running distances
    assert np.array_equal(furthest_point_sample(x, m).cpu().numpy(), ridx)


# ------------------------------------------------------------------ ball query
@pytest.mark.parametrize("n,m,r,ns", [(6000, 1500, 0.1, 32), (6000, 6000, 0.2, 32), (1500, 375, 0.4, 32),
                                      (375, 93, 0.8, 16), (3000, 700, 0.02, 8), (2000, 2000, 5.0, 32),
                                      (24000, 6000, 0.1, 32), (4096, 4096, 0.3, 64), (3000, 3000, 0.25, 100),
                                      (2500, 100, 10.0, 32)])
def test_ball_query(n, m, r, ns):
    from amcontrast3d_b200.layers import ball_query
    xyz, _ = scenes.batch_of_scenes(3, n, "surface", first_scene=4)
    q = np.ascontiguousarray(xyz[:, :m]) + (0.5 if r == 0.02 else 0.0)   # r=0.02 case: mostly empty balls
    ref = oo.ball_query(r, ns, xyz, q)
    idx = ball_query(r, ns, _t(xyz), _t(q))
    assert idx.dtype == torch.int32
    assert np.array_equal(idx.cpu().numpy(), ref)


def test_ball_query_same_tensor_and_lattice():
    """query set == support set (LocalAggregation): the culled path reuses the sorted support points as
    queries; lattice coordinates put many points at exactly d2 == r^2 (strict '<' must exclude them)."""
    from amcontrast3d_b200.layers import ball_query
    xyz, _ = scenes.batch_of_scenes(2, 6000, "surface", first_scene=9)
    t = _t(xyz)
    assert np.array_equal(ball_query(0.2, 32, t, t).cpu().numpy(), oo.ball_query(0.2, 32, xyz, xyz))
    lat = np.random.default_rng(1).integers(0, 24, size=(2, 5000, 3)).astype(np.float32) * 0.125
    tl = _t(lat)
    for r in (0.125, 0.25, 0.375):
        assert np.array_equal(ball_query(r, 16, tl, tl).cpu().numpy(), oo.ball_query(r, 16, lat, lat))


# ------------------------------------------------------------------ three_nn / interpolate
@pytest.mark.parametrize("n,m", [(375, 93), (1500, 375), (6000, 1500), (700, 2), (50, 1), (12000, 3000), (5000, 5000),
                                 (100, 2500)])
def test_three_nn(n, m):
    from amcontrast3d_b200.layers import three_nn
    xyz, _ = scenes.batch_of_scenes(2, max(n, m), "surface", first_scene=6)
    unknown, known = np.ascontiguousarray(xyz[:, :n]), np.ascontiguousarray(xyz[:, :m])
    rd2, ri = oo.three_nn(unknown, known)
    dist, idx = three_nn(_t(unknown), _t(known))
    assert np.array_equal(idx.cpu().numpy(), ri)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(rd2))


def test_three_nn_dense_queries_thread_per_query_path():
    """>= 128k query slots, at least twice as many queries as known points, k <= 4: the thread-per-query search
    (knn_grid.cu 4b).  Lattice points put exact distance ties on most rows."""
    from amcontrast3d_b200.layers import three_nn
    xyz, _ = scenes.batch_of_scenes(6, 22000, "surface", first_scene=16)
    known = np.ascontiguousarray(xyz[:, ::11])                               # 2000 known points per cloud
    rd2, ri = oo.three_nn(xyz, known)
    dist, idx = three_nn(_t(xyz), _t(known))
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), np.sqrt(rd2))
    rng = np.random.default_rng(4)
    lat_u = rng.integers(0, 24, size=(6, 22000, 3)).astype(np.float32) * 0.125
    lat_k = rng.integers(0, 24, size=(6, 1800, 3)).astype(np.float32) * 0.125
    rd2, ri = oo.three_nn(lat_u, lat_k)
    dist, idx = three_nn(_t(lat_u), _t(lat_k))
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), np.sqrt(rd2))


@pytest.mark.parametrize("k", [1, 2, 4])
def test_knn_dense_queries_thread_per_query_path(k):
    from amcontrast3d_b200 import _amloss
    xyz, _ = scenes.surface_scene(140000, seed=12)
    sup = np.ascontiguousarray(xyz[::14])                                    # 10 000 support points, one segment
    o, qo = np.array([sup.shape[0]], np.int32), np.array([xyz.shape[0]], np.int32)
    idx, d2 = _amloss.knn_raw(k, _t(sup), _t(xyz), _t(o), _t(qo))
    _check_knn(idx.cpu().numpy(), d2.cpu().numpy(), k, sup, xyz, o, qo, sqrt=False)


def test_knn_self_small_k_thread_per_query_path():
    """self search, k <= 4, >= 64k points: also served by the thread-per-query kernel"""
    from amcontrast3d_b200 import _amloss
    xyz, _ = scenes.volume_scene(66000, seed=21)
    o = np.array([66000], np.int32)
    idx, d2 = _amloss.knn_raw(3, _t(xyz), None, _t(o), _t(o))
    _check_knn(idx.cpu().numpy(), d2.cpu().numpy(), 3, xyz, xyz, o, o, sqrt=False)


def test_three_interpolate_forward_backward():
    from amcontrast3d_b200.layers import three_interpolate, three_interpolation
    rng = np.random.default_rng(2)
    B, C, m, n = 2, 96, 375, 1500
    xyz, _ = scenes.batch_of_scenes(B, n, "surface", first_scene=8)
    known = np.ascontiguousarray(xyz[:, :m])
    feats = rng.standard_normal((B, C, m)).astype(np.float32)
    rd2, ri = oo.three_nn(xyz, known)
    dist = np.sqrt(rd2)
    recip = (1.0 / (dist + np.float32(1e-8))).astype(np.float32)
    w = (recip / recip.sum(2, keepdims=True)).astype(np.float32)
    f = _t(feats).requires_grad_(True)
    out = three_interpolate(f, _t(ri), _t(w))
    assert np.array_equal(out.detach().cpu().numpy(), oo.three_interpolate(feats, ri, w))
    go = rng.standard_normal((B, C, n)).astype(np.float32)
    out.backward(_t(go))
    assert rel_err(f.grad.cpu().numpy(), oo.three_interpolate_grad(go, ri, w, m)) <= 1e-5
    # composed operator, weights formed by torch on device exactly as upsampling.py:97-100
    out2 = three_interpolation(_t(xyz), _t(known), _t(feats))
    assert np.allclose(out2.cpu().numpy(), out.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,C,m,n", [(1, 36, 50, 1001), (2, 100, 93, 375), (1, 8, 3, 64), (2, 130, 700, 2048),
                                     (1, 1024, 93, 372), (3, 6, 40, 100)])
def test_three_interpolate_odd_shapes(B, C, m, n):
    """channel counts that are not multiples of 32 / 4, output lengths that do and do not allow the bulk rows,
    fewer known points than neighbours-per-point variety: TMA-staged and direct kernels agree with the oracle"""
    from amcontrast3d_b200.layers import three_interpolate
    rng = np.random.default_rng(C + n)
    feats = rng.standard_normal((B, C, m)).astype(np.float32)
    idx = rng.integers(0, m, size=(B, n, 3)).astype(np.int32)
    w = rng.random((B, n, 3)).astype(np.float32)
    w /= w.sum(2, keepdims=True)
    f = _t(feats).requires_grad_(True)
    out = three_interpolate(f, _t(idx), _t(w))
    assert np.array_equal(out.detach().cpu().numpy(), oo.three_interpolate(feats, idx, w))
    go = rng.standard_normal((B, C, n)).astype(np.float32)
    out.backward(_t(go))
    assert rel_err(f.grad.cpu().numpy(), oo.three_interpolate_grad(go, idx, w, m)) <= 1e-5


# ------------------------------------------------------------------ grouping
@pytest.mark.parametrize("B,C,N,npoint,ns", [(2, 64, 6000, 1500, 32), (2, 128, 1500, 1500, 32), (3, 3, 6000, 1500, 32),
                                             (1, 40, 375, 93, 16), (2, 7, 500, 100, 5), (1, 1024, 93, 93, 32),
                                             (2, 100, 777, 301, 32), (1, 36, 50, 9, 64), (2, 12, 300, 77, 24),
                                             (1, 130, 1000, 1, 32), (2, 64, 40, 500, 8)])
def test_grouping_forward_backward(B, C, N, npoint, ns):
    from amcontrast3d_b200.layers import grouping_operation
    rng = np.random.default_rng(C)
    feats = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, npoint, ns)).astype(np.int32)
    idx[:, :, ns // 2:] = idx[:, :, :1]          # ball-query style padding: repeated first hit
    f = _t(feats).requires_grad_(True)
    out = grouping_operation(f, _t(idx))
    assert np.array_equal(out.detach().cpu().numpy(), oo.group_points(feats, idx))
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(_t(go))
    assert rel_err(f.grad.cpu().numpy(), oo.group_points_grad(go, idx, N)) <= 1e-5


def test_grouping_without_workspace_and_gather():
    from amcontrast3d_b200 import _capi
    from amcontrast3d_b200.layers import gather_operation
    rng = np.random.default_rng(3)
    B, C, N, npoint, ns = 2, 64, 2000, 500, 32
    feats = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, npoint, ns)).astype(np.int32)
    f, i = _t(feats), _t(idx)
    out = torch.empty((B, C, npoint, ns), device=DEV)
    _capi.call("amc3d_group_points", B, C, N, npoint, ns, f.data_ptr(), i.data_ptr(), out.data_ptr(), 0)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), oo.group_points(feats, idx))
    gidx = rng.integers(0, N, size=(B, 300)).astype(np.int32)
    fg = _t(feats).requires_grad_(True)
    g = gather_operation(fg, _t(gidx))
    assert np.array_equal(g.detach().cpu().numpy(), oo.gather_points(feats, gidx))
    go = rng.standard_normal(g.shape).astype(np.float32)
    g.backward(_t(go))
    assert rel_err(fg.grad.cpu().numpy(), oo.gather_points_grad(go, gidx, N)) <= 1e-5


def test_query_and_group_module():
    from amcontrast3d_b200.layers import create_grouper
    xyz, _ = scenes.batch_of_scenes(2, 3000, "surface", first_scene=12)
    q = np.ascontiguousarray(xyz[:, :750])
    feats = np.random.default_rng(0).standard_normal((2, 32, 3000)).astype(np.float32)
    grouper = create_grouper({"NAME": "ballquery", "radius": 0.15, "nsample": 32, "normalize_dp": True})
    dp, fj = grouper(_t(q), _t(xyz), _t(feats))
    idx = oo.ball_query(0.15, 32, xyz, q)
    ref_xyz = oo.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)
    ref_dp = (ref_xyz - q.transpose(0, 2, 1)[..., None]) / np.float32(0.15)
    # torch divides by the python scalar on device as x * (1/r): one ulp from numpy's true division
    assert np.allclose(dp.cpu().numpy(), ref_dp.astype(np.float32), rtol=3e-7, atol=0)
    assert np.array_equal(fj.cpu().numpy(), oo.group_points(feats, idx))


@pytest.mark.parametrize("relative,normalize,radius", [(True, True, 0.15), (True, True, 0.1), (True, True, 1.6),
                                                       (True, False, 0.2), (False, False, 0.2)])
def test_fused_relative_xyz_is_bit_identical_to_the_composition(relative, normalize, radius):
    """QueryAndGroup's one-kernel relative coordinates == transpose + grouping + subtraction + division as torch
    evaluates them on the device (group.py:244-249 of the reference), bit for bit."""
    from amcontrast3d_b200.layers import QueryAndGroup, ball_query, grouping_operation
    xyz, _ = scenes.batch_of_scenes(3, 4000, "surface", first_scene=15)
    sup, qry = _t(xyz), _t(np.ascontiguousarray(xyz[:, :1000]))
    grouper = QueryAndGroup(radius, 32, relative_xyz=relative, normalize_dp=normalize)
    dp, fj = grouper(qry, sup, None)
    assert fj is None
    idx = ball_query(radius, 32, sup, qry)
    ref = grouping_operation(sup.transpose(1, 2).contiguous(), idx)
    if relative:
        ref = ref - qry.transpose(1, 2).unsqueeze(-1)
        if normalize:
            ref /= radius
    assert torch.equal(dp, ref)
    # coordinates that require grad take the differentiable composition
    sup_g = sup.clone().requires_grad_(True)
    dp_g, _ = grouper(qry, sup_g, None)
    assert torch.equal(dp_g.detach(), ref)
    dp_g.sum().backward()
    assert sup_g.grad is not None


def test_pointops_grouping_packed_layout():
    from amcontrast3d_b200 import pointops
    rng = np.random.default_rng(5)
    n, c, m, ns = 3000, 64, 1000, 15
    feats = rng.standard_normal((n, c)).astype(np.float32)
    idx = rng.integers(0, n, size=(m, ns)).astype(np.int32)
    f = _t(feats).requires_grad_(True)
    out = pointops.grouping(f, _t(idx))
    assert np.array_equal(out.detach().cpu().numpy(), feats[idx])
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(_t(go))
    ref = np.zeros_like(feats)
    np.add.at(ref, idx.reshape(-1), go.reshape(-1, c))
    assert rel_err(f.grad.cpu().numpy(), ref) <= 1e-5
