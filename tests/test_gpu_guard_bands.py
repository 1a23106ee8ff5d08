"""Out-of-bounds WRITE check for the kernels added in round 2, without a sanitizer: every buffer the host code
hands to the library is carved out of a larger allocation filled with a sentinel byte pattern, and the bands on
both sides must be untouched after the call.  Ragged shapes on purpose (M not a multiple of the 128-row tile, O
off the 128-column slice, C off the 32-column K chunk, a batch of 3)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BAND = 4096          # bytes on each side
SENT = 0xA5


class _GuardedTorch:
    """stands in for the `torch` name inside a module: empty / empty_like return the middle of a sentinel-filled
    allocation, everything else is torch's"""

    def __init__(self):
        self.allocs = []

    def __getattr__(self, k):
        return getattr(torch, k)

    def empty(self, *size, dtype=None, device=None, **kw):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            size = tuple(size[0])
        dtype = dtype or torch.float32
        n = int(np.prod(size)) if len(size) else 1
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        raw = torch.full((BAND + nbytes + BAND,), SENT, dtype=torch.uint8, device=device)
        self.allocs.append((raw, nbytes))
        return raw[BAND:BAND + nbytes].view(dtype).view(size)

    def empty_like(self, t, **kw):
        return self.empty(tuple(t.shape), dtype=kw.get("dtype", t.dtype), device=t.device)

    def check(self):
        assert self.allocs
        for raw, nbytes in self.allocs:
            assert bool((raw[:BAND] == SENT).all()), "write below a buffer"
            assert bool((raw[BAND + nbytes:] == SENT).all()), "write above a buffer"
        return len(self.allocs)


@pytest.mark.parametrize("B,N,M,C,O,ns,prec", [(3, 333, 333, 40, 72, 16, "tf32x3"), (2, 500, 125, 64, 200, 32, "tf32"),
                                               (1, 260, 65, 264, 136, 32, "tf32x3"), (2, 129, 129, 8, 8, 16, "tf32"),
                                               # one output slice whose weights do NOT stay resident (C > 216), both precisions
                                               (1, 200, 200, 264, 72, 32, "tf32"), (1, 200, 200, 264, 72, 16, "tf32x3")])
def test_fused_operator_writes_stay_inside_its_buffers(monkeypatch, B, N, M, C, O, ns, prec):
    from amcontrast3d_b200 import scenes
    from amcontrast3d_b200.layers import ball_query, fused, _fused_backward
    g = _GuardedTorch()
    monkeypatch.setattr(fused, "torch", g)
    monkeypatch.setattr(_fused_backward, "torch", g)
    xyz, _ = scenes.batch_of_scenes(B, N, "surface", first_scene=5)
    p = torch.from_numpy(xyz).cuda()
    q = p[:, :M].contiguous()
    gen = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn(B, C, N, device="cuda", generator=gen).requires_grad_(True)
    w = (torch.randn(O, C + 3, device="cuda", generator=gen) * 0.2).requires_grad_(True)
    gam = torch.ones(O, device="cuda", requires_grad=True)
    bet = torch.zeros(O, device="cuda", requires_grad=True)
    idx = ball_query(0.25, ns, p, q)
    out, mean, var = fused.FusedGroupConvBNReLUMax.apply(f, w, gam, bet, q, p, idx, 0.25, True, 1e-5, prec)
    torch.cuda.synchronize()
    n_fwd = g.check()
    assert n_fwd >= 9
    out.backward(torch.randn(out.shape, device="cuda", generator=gen))
    torch.cuda.synchronize()
    assert g.check() > n_fwd
    assert torch.isfinite(out).all() and torch.isfinite(f.grad).all() and torch.isfinite(w.grad).all()


def test_voxel_and_crop_kernels_write_inside_their_buffers(monkeypatch):
    from amcontrast3d_b200 import data_util as du
    g = _GuardedTorch()
    monkeypatch.setattr(du, "torch", g)
    gen = torch.Generator(device="cuda").manual_seed(2)
    for n in (1, 777, 100003):
        c = torch.rand(n, 3, device="cuda", generator=gen) * 7
        du.voxelize(c, 0.04, mode=1)
        du.crop_pc(c, None, torch.zeros(n, dtype=torch.long, device="cuda"), split="val", voxel_max=max(1, n // 3))
        torch.cuda.synchronize()
    g.check()


def test_class_counts_writes_inside_its_buffer():
    from amcontrast3d_b200 import _capi
    from amcontrast3d_b200._capi import ptr, stream
    g = _GuardedTorch()
    m, ncls, k = 5001, 13, 16
    gen = torch.Generator(device="cuda").manual_seed(3)
    tgt = torch.randint(0, ncls, (m,), device="cuda", generator=gen, dtype=torch.int32)
    pred = torch.randint(0, ncls, (m,), device="cuda", generator=gen, dtype=torch.int32)
    nbr = torch.randint(0, m, (m, k), device="cuda", generator=gen, dtype=torch.int32)
    out = g.empty((3 * ncls,), dtype=torch.float32, device="cuda")
    with _capi.guard(tgt):
        _capi.call("amc3d_class_counts", m, ncls, ptr(tgt), ptr(pred), ptr(nbr), k, ptr(out), stream(tgt))
    torch.cuda.synchronize()
    g.check()
    t, pr = tgt.cpu().numpy(), pred.cpu().numpy()
    want = np.concatenate([np.bincount(t[t == pr], minlength=ncls), np.bincount(pr, minlength=ncls),
                           np.bincount(t, minlength=ncls)])
    assert np.array_equal(out.cpu().numpy(), want.astype(np.float32))
