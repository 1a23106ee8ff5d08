"""Tensor-level wrappers of the AM-loss / refinement entry points of include/amc3d.h.

Everything here runs on the CUDA device of its inputs through libamc3d; there is no CPU path.
The modules under amcontrast3d_b200/AMContrast3D/ (which mirror the reference's classes) are
thin compositions of these functions.
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function

from . import _capi
from ._capi import LossParams, ptr, stream

_CCTYPE = {"Method1": 1, "Method2": 2, "Method3": 3}
KE_MAX = 32  # neighbours per point the fused kernels keep in registers (amloss.cu KE_MAX)


def _i32(t):
    return t if t.dtype == torch.int32 else t.int()


def knn_raw(nsample: int, xyz, new_xyz, offset, new_offset, want_order: bool = False):
    """amc3d_knnquery -> (idx (m,nsample) i32, dist2 (m,nsample) f32 SQUARED); no sqrt, no autograd.
    want_order=True additionally returns the spatially coherent query permutation (m) i32 of
    amc3d_knnquery_order, which the fused loss uses as its anchor order."""
    if new_xyz is None:
        new_xyz = xyz
    assert xyz.is_contiguous() and new_xyz.is_contiguous()
    nsample = int(nsample)
    m = new_xyz.shape[0]
    idx = torch.empty((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.empty((m, nsample), dtype=torch.float32, device=xyz.device)
    offset, new_offset = _i32(offset).contiguous(), _i32(new_offset).contiguous()
    order = torch.empty((m,), dtype=torch.int32, device=xyz.device) if want_order else None
    with _capi.guard(xyz):
        _capi.call("amc3d_knnquery_order", int(xyz.shape[0]), m, int(offset.shape[0]), nsample, ptr(xyz), ptr(new_xyz),
                   ptr(offset), ptr(new_offset), ptr(idx), ptr(dist2), ptr(order), stream(xyz))
    if want_order:
        return idx, dist2, order
    return idx, dist2


def stage_labels(target, num_classes: int, ignore_index, nidx=None):
    """Integer stage labels (AEF/utils.py:11-43 + the argmax of MarginContrast.py:112).
    target (M0) i64; nidx (m,kr) i32 kNN rows into the stage-0 points, or None for stage 0.
    Returns (cls (m) i32, ncls) with ncls = num_classes (+1 when ignore_index is not None)."""
    has_ignore = ignore_index is not None
    ncls = num_classes + (1 if has_ignore else 0)
    target = target.contiguous()
    assert target.dtype == torch.int64
    if nidx is None:
        m, kr = target.shape[0], 0
    else:
        assert nidx.is_contiguous() and nidx.dtype == torch.int32
        m, kr = nidx.shape
    cls = torch.empty((m,), dtype=torch.int32, device=target.device)
    with _capi.guard(target):
        _capi.call("amc3d_stage_labels", m, kr, ncls, int(has_ignore), int(ignore_index) if has_ignore else 0,
                   ptr(target), ptr(nidx), ptr(cls), stream(target))
    return cls, ncls


class NeighbourList:
    """(nbr pointer, ld, ke) view of the non-self columns of a kNN result, without a copy."""

    def __init__(self, idx: torch.Tensor, drop_self: bool):
        assert idx.is_contiguous() and idx.dtype == torch.int32 and idx.dim() == 2
        self.tensor = idx  # keeps the storage alive
        self.m, k = idx.shape
        self.ld = k
        self.ke = k - 1 if drop_self else k
        self.ptr = idx.data_ptr() + (4 if drop_self else 0)
        if self.ke < 1 or self.ke > KE_MAX:
            raise _capi.Amc3dError(f"the fused AM-loss kernels support 1..{KE_MAX} neighbours per point, got {self.ke}")


def posmask_count(nl: NeighbourList, cls):
    """-> (posbits (m) i32 bit j = same class as neighbour j, cnt (m) i32, max_cnt (1) i32)"""
    dev = cls.device
    posbits = torch.empty((nl.m,), dtype=torch.int32, device=dev)
    cnt = torch.empty((nl.m,), dtype=torch.int32, device=dev)
    max_cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    with _capi.guard(cls):
        _capi.call("amc3d_posmask_count", nl.m, nl.ke, nl.ld, nl.ptr, ptr(cls), ptr(posbits), ptr(cnt),
                   ptr(max_cnt), stream(cls))
    return posbits, cnt, max_cnt


# Which torch backend's FP32 rounding of the reference's square_distance the ambiguity kernel reproduces
# (amc3d.h, amc3d_ambiguity_backend): "cuda" — what the reference computes where it actually runs (its
# ambiguity.py hard-codes .cuda()) — or "cpu" — what the reference's modules give when run on CPU, as for
# tests/golden/loss_golden.npz.  Env AMC3D_AMBIGUITY_BACKEND sets the initial value.
import os as _os
AMBIGUITY_BACKEND = _os.environ.get("AMC3D_AMBIGUITY_BACKEND", "cuda")


class ambiguity_backend:
    """with ambiguity_backend("cpu"): ...   (tests against the CPU-generated golden vectors)"""

    def __init__(self, name):
        assert name in ("cpu", "cuda")
        self.name = name

    def __enter__(self):
        global AMBIGUITY_BACKEND
        self.prev, AMBIGUITY_BACKEND = AMBIGUITY_BACKEND, self.name

    def __exit__(self, *exc):
        global AMBIGUITY_BACKEND
        AMBIGUITY_BACKEND = self.prev
        return False


def ambiguity(p, nl: NeighbourList, posbits, cnt, max_cnt, cctype: str, beta: float, nu: float):
    """-> (a (m) f32, stats (8) i32: [#selected, #boundary, 5 report bins, 0])"""
    assert p.is_contiguous() and p.dtype == torch.float32
    a = torch.empty((nl.m,), dtype=torch.float32, device=p.device)
    stats = torch.zeros((8,), dtype=torch.int32, device=p.device)
    with _capi.guard(p):
        _capi.call("amc3d_ambiguity_backend", nl.m, nl.ke, nl.ld, ptr(p), nl.ptr, ptr(posbits), ptr(cnt), ptr(max_cnt),
                   _CCTYPE[cctype], float(beta), float(nu), {"cpu": 0, "cuda": 1}[AMBIGUITY_BACKEND], ptr(a),
                   ptr(stats), stream(p))
    return a, stats


def unpack_posmask(posbits, ke: int):
    """(m) i32 bitmask -> (m,ke) bool, for API functions that return the reference's posmask"""
    shifts = torch.arange(ke, device=posbits.device, dtype=torch.int32)
    return ((posbits.unsqueeze(-1) >> shifts) & 1).bool()


def pack_posmask(posmask):
    """(m,ke) bool -> ((m) i32 bitmask, (m) i32 count)"""
    ke = posmask.shape[1]
    w = (1 << torch.arange(ke, device=posmask.device, dtype=torch.int64))
    bits = (posmask.long() * w).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).int()
    return bits, posmask.sum(-1).int()


def loss_params(args) -> LossParams:
    """ambiguity_args -> amc3d_loss_params; raises for the combinations only torch composes"""
    t = getattr(args, "temperature", None)
    margin = {"constant": 0, "adaptive": 1}[args.margin]
    db = {"-m": 1, "+m": 2}.get(args.db, 0)
    cl = {"Method1": 1, "Method2": 2}[args.supervisedCL]
    return LossParams(float(t) if t is not None else 1.0, int(t is not None), margin, float(getattr(args, "mu", 0.0)),
                      float(args.nu), db, cl)


def fused_supported(args) -> bool:
    return args.margin in ("constant", "adaptive") and args.supervisedCL in ("Method1", "Method2")


class AMLossFunction(Function):
    """L_s = mean_{i: 0<a_i<=1} -log r_i for one stage (MarginContrast.py:250-257), fused.

    forward launches row_inv_norm, amloss_forward (loss rows + dL/du accumulation) and the
    deterministic reduction; backward launches amloss_backward, which reads the upstream
    scalar and |sel| from device memory — no host synchronisation anywhere."""

    @staticmethod
    def forward(ctx, f, nl: NeighbourList, posbits, a, stats, params: LossParams, order=None):
        assert f.is_contiguous() and f.dtype == torch.float32 and f.dim() == 2
        m, d = f.shape
        dev = f.device
        inv = torch.empty((m,), dtype=torch.float32, device=dev)
        # zero-filled: with a compacted `order` (see compact_order) the kernel never visits unselected anchors
        loss_pt = torch.zeros((m,), dtype=torch.float32, device=dev)
        ghat = torch.zeros((m, d), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with _capi.guard(f):
            st = stream(f)
            _capi.call("amc3d_row_inv_norm", m, d, ptr(f), ptr(inv), st)
            _capi.call("amc3d_amloss_forward_order", m, d, nl.ke, nl.ld, ptr(f), ptr(inv), nl.ptr, ptr(posbits),
                       ptr(a), ctypes.byref(params), ptr(loss_pt), ptr(ghat), ptr(order), st)
            _capi.call("amc3d_amloss_reduce", m, ptr(loss_pt), ptr(stats), ptr(loss), st)
        ctx.save_for_backward(f, inv, ghat, stats)
        ctx.loss_rows = loss_pt
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        f, inv, ghat, stats = ctx.saved_tensors
        m, d = f.shape
        up = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        grad_f = torch.empty_like(f)
        with _capi.guard(f):
            _capi.call("amc3d_amloss_backward", m, d, ptr(f), ptr(inv), ptr(ghat), ptr(up), ptr(stats), 0,
                       ptr(grad_f), stream(f))
        return grad_f, None, None, None, None, None, None


def compact_order(order, a):
    """The visiting order with the anchors the loss does not select (a outside (0, 1]) removed: selected
    anchors first, in their original relative order, then -1 up to the original length.  The forward kernel
    skips negative entries, so its warps are full of selected anchors instead of being kept alive by one
    selected anchor among several (43 % of the stage-0 anchors are selected at config 2, 67 % of the warps
    had at least one).  Static shapes, no host synchronisation: capturable."""
    m = order.numel()
    keep = torch.logical_and(a > 0, a <= 1)[order.long()]
    slot = torch.cumsum(keep, 0) - 1
    slot = torch.where(keep, slot, torch.full_like(slot, m))          # dropped anchors go to a dump slot
    out = torch.full((m + 1,), -1, dtype=torch.int32, device=order.device)
    out.scatter_(0, slot, order)
    return out[:m]


def am_loss(f, nl, posbits, a, stats, args, order=None):
    f = f.contiguous()
    if f.dtype != torch.float32:
        f = f.float()
    return AMLossFunction.apply(f, nl, posbits, a, stats, loss_params(args), order)


# --------------------------------------------------------------------------------------------
# masked refinement
# --------------------------------------------------------------------------------------------
def refine_select(nl: NeighbourList, a_flat):
    jmin = torch.empty((nl.m,), dtype=torch.int32, device=a_flat.device)
    with _capi.guard(a_flat):
        _capi.call("amc3d_refine_select", nl.m, nl.ke, nl.ld, nl.ptr, ptr(a_flat), ptr(jmin), stream(a_flat))
    return jmin


class DualMasksFunction(Function):
    """out = gamma*(f*~mask + chunk[jmin]*mask) + (1-gamma)*f over a contiguous (B,D,n) buffer
    (MaskedRefine.py:62-81).  Gradient flows to f only."""

    @staticmethod
    def forward(ctx, f, a, jmin, thr, thr_max, gamma, count):
        B, D, n = f.shape
        out = torch.empty_like(f)
        with _capi.guard(f):
            _capi.call("amc3d_refine_forward", B, D, n, ptr(f), ptr(a), ptr(jmin), float(thr), float(thr_max),
                       float(gamma), ptr(out), ptr(count), stream(f))
        ctx.save_for_backward(a, jmin)
        ctx.cfg = (float(thr), float(thr_max), float(gamma))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        a, jmin = ctx.saved_tensors
        thr, thr_max, gamma = ctx.cfg
        go = grad_out.contiguous()
        B, D, n = go.shape
        grad_f = torch.empty_like(go)
        with _capi.guard(go):
            _capi.call("amc3d_refine_backward", B, D, n, ptr(go), ptr(a), ptr(jmin), thr, thr_max, gamma,
                       ptr(grad_f), stream(go))
        return grad_f, None, None, None, None, None, None
