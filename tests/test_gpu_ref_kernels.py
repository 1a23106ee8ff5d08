"""The sm_100a kernels against the REFERENCE's own CUDA kernels (oracle/_ref, compiled unmodified
from the reference sources for sm_100) on the same device and inputs."""
import numpy as np
import pytest
import torch

from amcontrast3d_b200 import scenes
from oracle import ref_kernels as rk

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rk.available(), reason="oracle/_ref not built")]
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_knn_vs_reference_kernel():
    from amcontrast3d_b200 import _amloss
    xyz, _ = scenes.batch_of_scenes(2, 12000, "surface", first_scene=50)
    flat = _t(xyz.reshape(-1, 3))
    o = _t(np.array([24000], dtype=np.int32))
    for k in (4, 16, 24, 32, 64):
        ri, rd = rk.knnquery(k, flat, flat, o, o)
        _, rdp = rk.knnquery(k + 1, flat, flat, o, o)
        i, d = _amloss.knn_raw(k, flat, flat, o, o)
        assert torch.equal(d, rd)                            # sorted distances: identical on every row
        tie_free = (rdp[:, 1:] > rdp[:, :-1]).all(1)         # rows whose k+1 nearest distances are distinct
        assert tie_free.float().mean() > 0.98
        assert torch.equal(i[tie_free], ri[tie_free])


def test_knn_small_k_vs_reference_kernel():
    """k <= 4 with >= 64k query slots goes through the thread-per-query search (self and dense queries)"""
    from amcontrast3d_b200 import _amloss
    xyz, _ = scenes.batch_of_scenes(4, 24000, "surface", first_scene=55)
    flat = _t(xyz.reshape(-1, 3))
    o = _t(np.array([96000], dtype=np.int32))
    sup = flat[::6].contiguous()
    so = _t(np.array([sup.shape[0]], dtype=np.int32))
    for k in (1, 2, 3, 4):
        for s_xyz, s_off in ((flat, o), (sup, so)):
            ri, rd = rk.knnquery(k, s_xyz, flat, s_off, o)
            _, rdp = rk.knnquery(k + 1, s_xyz, flat, s_off, o)
            i, d = _amloss.knn_raw(k, s_xyz, flat, s_off, o)
            assert torch.equal(d, rd)
            tie_free = (rdp[:, 1:] > rdp[:, :-1]).all(1)
            assert tie_free.float().mean() > 0.98
            assert torch.equal(i[tie_free], ri[tie_free])


def test_fps_ball_three_nn_vs_reference_kernels():
    from amcontrast3d_b200.layers import ball_query, furthest_point_sample, three_nn
    xyz, _ = scenes.batch_of_scenes(4, 24000, "surface", first_scene=60)
    p = _t(xyz)
    ridx, _ = rk.fps(p, 6000)
    idx = furthest_point_sample(p, 6000)
    assert torch.equal(idx, ridx)
    q = torch.gather(p, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    assert torch.equal(ball_query(0.1, 32, p, q), rk.ball_query(0.1, 32, p, q))
    assert torch.equal(ball_query(0.2, 32, q, q), rk.ball_query(0.2, 32, q, q))
    rd2, ri = rk.three_nn(p, q)
    dist, i3 = three_nn(p, q)
    assert torch.equal(i3, ri) and torch.equal(dist, torch.sqrt(rd2))
    # lattice: exact ties through the reference's tree reduction
    lat = _t(np.random.default_rng(0).integers(0, 16, size=(2, 6000, 3)).astype(np.float32) * 0.125)
    assert torch.equal(furthest_point_sample(lat, 1500), rk.fps(lat, 1500)[0])


def test_grouping_and_interpolate_vs_reference_kernels():
    from amcontrast3d_b200.layers import grouping_operation, three_interpolate
    rng = np.random.default_rng(1)
    B, C, N, npnt, ns = 2, 128, 6000, 6000, 32
    f = _t(rng.standard_normal((B, C, N)).astype(np.float32)).requires_grad_(True)
    idx = _t(rng.integers(0, N, size=(B, npnt, ns)).astype(np.int32))
    out = grouping_operation(f, idx)
    assert torch.equal(out.detach(), rk.group_points(f.detach(), idx))
    go = _t(rng.standard_normal(tuple(out.shape)).astype(np.float32))
    out.backward(go)
    ref = rk.group_points_grad(go, idx, N)
    assert (f.grad - ref).norm() <= 1e-5 * ref.norm()
    i3 = _t(rng.integers(0, N, size=(B, 1500, 3)).astype(np.int32))
    w = _t(rng.random((B, 1500, 3)).astype(np.float32))
    assert torch.equal(three_interpolate(f.detach(), i3, w), rk.three_interpolate(f.detach(), i3, w))
