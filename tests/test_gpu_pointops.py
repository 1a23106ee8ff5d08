"""GPU parity of the remaining pointops operators (furthestsampling, ballquery, interpolation, subtraction,
aggregation on the packed layout) against the CPU restatement of the reference kernels (oracle_pop_* in
oracle/ops_oracle.c) and, where oracle/_ref is built, against the reference's own CUDA kernels on the same
device.  Called through the reference-facing functions of amcontrast3d_b200.pointops, i.e. through the
C-ABI.  Bar: identical indices, gathers and FMA-chain outputs; atomically accumulated gradients within 1e-5
relative (the reference's own atomics are order-dependent)."""
import numpy as np
import pytest
import torch

from _util import rel_err
from amcontrast3d_b200 import pointops, scenes
from oracle import ops_oracle as oo
from oracle import ref_kernels as rk

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _n(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("sizes,samples", [
    ([1200, 300, 2500], [300, 40, 560]),          # ragged
    ([2048, 2048, 2048, 2048], [512] * 4),        # equal segments: one batched cluster launch
    ([5, 1, 700], [5, 1, 100]),                   # tiny segments, m == n
    ([24000], [6000]),
])
def test_furthestsampling(sizes, samples):
    rng = np.random.default_rng(len(sizes))
    xyz = rng.random((sum(sizes), 3)).astype(np.float32)
    off, noff = np.cumsum(sizes).astype(np.int32), np.cumsum(samples).astype(np.int32)
    idx = pointops.furthestsampling(_t(xyz), _t(off), _t(noff))
    assert idx.dtype == torch.int32 and idx.shape == (int(noff[-1]),)
    assert np.array_equal(_n(idx), oo.pop_furthestsampling(xyz, off, noff))


def test_furthestsampling_exact_ties():
    rng = np.random.default_rng(3)
    lat = rng.integers(0, 8, size=(5000, 3)).astype(np.float32) * 0.25
    off, noff = np.array([1000, 3000, 5000], np.int32), np.array([250, 750, 1250], np.int32)
    idx = pointops.furthestsampling(_t(lat), _t(off), _t(noff))
    assert np.array_equal(_n(idx), oo.pop_furthestsampling(lat, off, noff))


@pytest.mark.parametrize("radius,nsample", [(0.1, 16), (0.3, 8), (0.01, 4), (0.2, 64)])
def test_ballquery_segments(radius, nsample):
    xyz, _ = scenes.batch_of_scenes(3, 1500, "surface", first_scene=80)
    xyz = np.ascontiguousarray(xyz.reshape(-1, 3))[:4100]
    off = np.array([1500, 1700, 4100], np.int32)
    qoff = np.array([400, 450, 1000], np.int32)
    q = np.concatenate([xyz[:400], xyz[1500:1550], xyz[1700:2250]])
    idx = pointops.ballquery(radius, nsample, _t(xyz), _t(q), _t(off), _t(qoff))
    assert np.array_equal(_n(idx), oo.pop_ballquery(radius, nsample, xyz, q, off, qoff))


def test_ballquery_single_segment_culled_and_self():
    xyz, _ = scenes.surface_scene(20000, seed=9)
    o = np.array([20000], np.int32)
    for r, ns in ((0.08, 32), (0.02, 16)):
        idx = pointops.ballquery(r, ns, _t(xyz), None, _t(o), _t(o))       # new_xyz=None -> xyz
        assert np.array_equal(_n(idx), oo.pop_ballquery(r, ns, xyz, None, o, o))
    q = np.ascontiguousarray(xyz[::7] + np.float32(0.3))                    # many queries without any hit
    qo = np.array([q.shape[0]], np.int32)
    idx = pointops.ballquery(0.05, 8, _t(xyz), _t(q), _t(o), _t(qo))
    assert np.array_equal(_n(idx), oo.pop_ballquery(0.05, 8, xyz, q, o, qo))


def _case(seed, n, ns, c, w_c):
    rng = np.random.default_rng(seed)
    return dict(a=rng.standard_normal((n, c)).astype(np.float32), b=rng.standard_normal((n, c)).astype(np.float32),
                idx=rng.integers(0, n, size=(n, ns)).astype(np.int32),
                pos=rng.standard_normal((n, ns, c)).astype(np.float32),
                w=rng.standard_normal((n, ns, w_c)).astype(np.float32),
                g=rng.standard_normal((n, c)).astype(np.float32),
                g3=rng.standard_normal((n, ns, c)).astype(np.float32))


@pytest.mark.parametrize("n,ns,c,w_c", [(700, 8, 32, 8), (513, 5, 7, 7), (2000, 16, 64, 8), (1, 1, 1, 1)])
def test_subtraction_and_aggregation(n, ns, c, w_c):
    d = _case(n, n, ns, c, w_c)
    a, b, pos, w = (_t(d[k]).requires_grad_(True) for k in ("a", "b", "pos", "w"))
    idx = _t(d["idx"])
    out = pointops.subtraction(a, b, idx)
    assert np.array_equal(_n(out), oo.pop_subtraction_fwd(d["a"], d["b"], d["idx"]))
    out.backward(_t(d["g3"]))
    e1, e2 = oo.pop_subtraction_bwd(d["idx"], d["g3"])
    assert rel_err(_n(a.grad), e1) < 1e-5 and rel_err(_n(b.grad), e2) < 1e-5
    a.grad = None
    out = pointops.aggregation(a, pos, w, idx)
    assert np.array_equal(_n(out), oo.pop_aggregation_fwd(d["a"], d["pos"], d["w"], d["idx"]))
    out.backward(_t(d["g"]))
    ei, ep, ew = oo.pop_aggregation_bwd(d["a"], d["pos"], d["w"], d["idx"], d["g"])
    assert np.array_equal(_n(pos.grad), ep)
    assert rel_err(_n(a.grad), ei) < 1e-5 and rel_err(_n(w.grad), ew) < 1e-5


def test_interpolation_kernels_and_both_front_ends():
    from amcontrast3d_b200 import pointops_cuda
    rng = np.random.default_rng(5)
    m, n, c, k = 900, 2600, 48, 3
    src = rng.standard_normal((m, c)).astype(np.float32)
    idx = rng.integers(0, m, size=(n, k)).astype(np.int32)
    w = rng.random((n, k)).astype(np.float32)
    init = rng.standard_normal((n, c)).astype(np.float32)
    out = _t(init)
    pointops_cuda.interpolation_forward_cuda(n, c, k, _t(src), _t(idx), _t(w), out)     # accumulates
    assert np.array_equal(_n(out), oo.pop_interpolation_fwd(src, idx, w, output=init))
    g = rng.standard_normal((n, c)).astype(np.float32)
    gi = torch.zeros((m, c), device=DEV)
    pointops_cuda.interpolation_backward_cuda(n, c, k, _t(g), _t(idx), _t(w), gi)
    assert rel_err(_n(gi), oo.pop_interpolation_bwd(g, idx, w, m)) < 1e-5
    # interpolation (torch composition) and interpolation2 (fused kernels) agree, forward and gradient
    xyz, _ = scenes.surface_scene(m, seed=1)
    new_xyz, _ = scenes.surface_scene(n, seed=2)
    o, no = _t(np.array([m], np.int32)), _t(np.array([n], np.int32))
    f1 = _t(src).requires_grad_(True)
    f2 = _t(src).requires_grad_(True)
    y1 = pointops.interpolation(_t(xyz), _t(new_xyz), f1, o, no)
    y2 = pointops.interpolation2(_t(xyz), _t(new_xyz), f2, o, no)
    assert rel_err(_n(y2), _n(y1)) < 1e-6
    y1.backward(_t(g))
    y2.backward(_t(g))
    assert rel_err(_n(f2.grad), _n(f1.grad)) < 1e-5


def test_querygroup_and_queryandgroup():
    xyz, _ = scenes.surface_scene(3000, seed=4)
    rng = np.random.default_rng(6)
    feat = rng.standard_normal((3000, 16)).astype(np.float32)
    o = np.array([3000], np.int32)
    q = np.ascontiguousarray(xyz[::4])
    qo = np.array([q.shape[0]], np.int32)
    gx, gf = pointops.querygroup(8, _t(xyz), _t(q), _t(feat), _t(o), _t(qo))
    ki, _ = oo.knnquery(8, xyz, q, o, qo, lex=True)
    assert np.array_equal(_n(gx), xyz[ki.astype(np.int64)] - q[:, None, :])
    assert np.array_equal(_n(gf), feat[ki.astype(np.int64)])
    gx, gf = pointops.querygroup(8, _t(xyz), _t(q), _t(feat), _t(o), _t(qo), radius=0.1, query_method="ball",
                                 normalize_dp=True)
    bi = oo.pop_ballquery(0.1, 8, xyz, q, o, qo)
    # torch divides a CUDA tensor by a host scalar as a multiplication by its float32 reciprocal
    assert np.array_equal(_n(gx), (xyz[bi.astype(np.int64)] - q[:, None, :]) * (np.float32(1) / np.float32(0.1)))
    both = pointops.queryandgroup(8, _t(xyz), _t(q), _t(feat), None, _t(o), _t(qo))
    assert both.shape == (q.shape[0], 8, 19)
    assert np.array_equal(_n(both[..., 3:]), feat[ki.astype(np.int64)])


@pytest.mark.skipif(not rk.pointops_available(), reason="oracle/_ref/libref_pointops.so not built")
def test_against_reference_kernels():
    from amcontrast3d_b200 import pointops_cuda
    xyz, _ = scenes.batch_of_scenes(3, 8000, "surface", first_scene=90)
    xyz = _t(xyz.reshape(-1, 3))
    off = _t(np.array([8000, 16000, 24000], np.int32))
    noff = _t(np.array([2000, 4000, 6000], np.int32))
    idx = pointops.furthestsampling(xyz, off, noff)
    assert torch.equal(idx, rk.pop_furthestsampling(xyz, off, noff)[0])
    roff = _t(np.array([5000, 5100, 24000], np.int32))          # ragged: 5000 / 100 / 18900
    rnoff = _t(np.array([1000, 1100, 3000], np.int32))
    assert torch.equal(pointops.furthestsampling(xyz, roff, rnoff), rk.pop_furthestsampling(xyz, roff, rnoff)[0])
    q = xyz[idx.long()].contiguous()
    for r, ns in ((0.1, 32), (0.03, 8)):
        assert torch.equal(pointops.ballquery(r, ns, xyz, q, off, noff), rk.pop_ballquery(r, ns, xyz, q, off, noff))
    one, qone = _t(np.array([24000], np.int32)), _t(np.array([6000], np.int32))
    assert torch.equal(pointops.ballquery(0.1, 32, xyz, q, one, qone), rk.pop_ballquery(0.1, 32, xyz, q, one, qone))
    d = _case(11, 4000, 16, 64, 8)
    a, b, pos, w, g, g3 = (_t(d[k]) for k in ("a", "b", "pos", "w", "g", "g3"))
    nidx = _t(d["idx"])
    assert torch.equal(pointops.subtraction(a, b, nidx), rk.pop_subtraction_fwd(a, b, nidx))
    assert torch.equal(pointops.aggregation(a, pos, w, nidx), rk.pop_aggregation_fwd(a, pos, w, nidx))
    gi, gp, gw = (torch.zeros_like(t) for t in (a, pos, w))
    pointops_cuda.aggregation_backward_cuda(4000, 16, 64, 8, a, pos, w, nidx, g, gi, gp, gw)
    ri, rp, rw = rk.pop_aggregation_bwd(a, pos, w, nidx, g)
    assert torch.equal(gp, rp)
    assert (gi - ri).norm() <= 1e-5 * ri.norm() and (gw - rw).norm() <= 1e-5 * rw.norm()
    iw = torch.rand((4000, 3), device=DEV)
    iidx = nidx[:, :3].contiguous()
    out = torch.zeros((4000, 64), device=DEV)
    pointops_cuda.interpolation_forward_cuda(4000, 64, 3, a, iidx, iw, out)
    assert torch.equal(out, rk.pop_interpolation_fwd(a, iidx, iw))
