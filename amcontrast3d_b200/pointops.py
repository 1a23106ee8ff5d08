"""Mirror of openpoints/cpp/pointops/functions/pointops.py: the operators on packed (n,3) / (n,c) tensors
with cumulative i32 offsets.

``knnquery(nsample, xyz, new_xyz, offset, new_offset) -> (idx i32 (m,nsample), dist f32)``
(pointops.py:32-56) is what AMContrast3D calls.  The rest of that module's public names —
furthestsampling, ballquery, grouping, querygroup, queryandgroup, subtraction, aggregation,
interpolation, interpolation2 — are here with the same signatures and results.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import pointops_cuda


class KNNQuery(Function):
    @staticmethod
    def forward(ctx, nsample, xyz, new_xyz, offset, new_offset):
        """xyz (n,3), new_xyz (m,3) or None (= xyz), offset/new_offset (b) i32 cumulative ends.
        `nsample` may be an int or a 0-dim tensor (AEF/utils.py:29 passes torch.prod(...))."""
        if new_xyz is None:
            new_xyz = xyz
        assert xyz.is_contiguous() and new_xyz.is_contiguous()
        nsample = int(nsample)
        m = new_xyz.shape[0]
        idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
        dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
        pointops_cuda.knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(idx, dist)
        return idx, dist

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None, None, None, None


knnquery = KNNQuery.apply


class Grouping(Function):
    @staticmethod
    def forward(ctx, input, idx):
        """input (n,c), idx (m,nsample) -> (m,nsample,c)"""
        assert input.is_contiguous() and idx.is_contiguous()
        m, nsample, n, c = idx.shape[0], idx.shape[1], input.shape[0], input.shape[1]
        output = torch.empty((m, nsample, c), dtype=torch.float32, device=input.device)
        pointops_cuda.grouping_forward_cuda(m, nsample, c, input, idx, output)
        ctx.n = n
        ctx.save_for_backward(idx)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        n = ctx.n
        idx, = ctx.saved_tensors
        m, nsample, c = grad_output.shape
        grad_input = torch.zeros((n, c), dtype=torch.float32, device=grad_output.device)
        pointops_cuda.grouping_backward_cuda(m, nsample, c, grad_output.contiguous(), idx, grad_input)
        return grad_input, None


grouping = Grouping.apply


def _buf(shape, like, dtype=torch.float32, zero=True):
    make = torch.zeros if zero else torch.empty
    return make(shape, dtype=dtype, device=like.device)


class FurthestSampling(Function):
    """xyz (n,3), offset (b), new_offset (b) -> idx (new_offset[-1]) i32: per segment, the furthest-point
    sample of new_offset[s]-new_offset[s-1] points as global indices (pointops.py:10-28)."""

    @staticmethod
    def forward(ctx, xyz, offset, new_offset):
        assert xyz.is_contiguous()
        h_off = offset.detach().to("cpu", torch.int64)
        sizes = torch.diff(h_off, prepend=h_off.new_zeros(1))
        n_max = int(sizes.max()) if sizes.numel() else 0
        total = int(new_offset[-1]) if new_offset.numel() else 0
        idx = _buf((total,), xyz, torch.int32)
        tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
        pointops_cuda.furthestsampling_cuda(offset.shape[0], n_max, xyz, offset, new_offset, tmp, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, g=None):
        return None, None, None


furthestsampling = FurthestSampling.apply


class BallQuery(Function):
    """radius, nsample, xyz (n,3), new_xyz (m,3) or None, offsets -> idx (m,nsample) i32 (pointops.py:59-76)"""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz, offset, new_offset):
        if new_xyz is None:
            new_xyz = xyz
        assert xyz.is_contiguous() and new_xyz.is_contiguous()
        idx = _buf((new_xyz.shape[0], int(nsample)), xyz, torch.int32)
        pointops_cuda.ballquery_cuda(new_xyz.shape[0], radius, int(nsample), xyz, new_xyz, offset, new_offset, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, g=None):
        return None, None, None, None, None, None


ballquery = BallQuery.apply


def _gather_rows(table, idx):
    """table (n,c), idx (m,nsample) i32 -> (m,nsample,c) by torch indexing (differentiable w.r.t. table)"""
    m, nsample = idx.shape
    return table[idx.reshape(-1).long(), :].view(m, nsample, table.shape[1])


def _offsets_from(xyz, new_xyz, idx):
    """neighbour coordinates relative to their query, (m,nsample,3)"""
    rel = _gather_rows(xyz, idx)
    rel -= new_xyz.unsqueeze(1)
    return rel


def querygroup(nsample, xyz, new_xyz, feat, offset, new_offset, radius=None, query_method='knn',
               normalize_dp=False, idx=None):
    """Neighbour search (kNN or ball) followed by grouping of coordinates (relative to the query) and
    features -> (grouped_xyz (m,nsample,3), grouped_feat (m,nsample,c) or None)  (pointops.py:111-158).

    query_method 'knn' / 'knnquery' selects the kNN, anything else the ball query of `radius`.  With
    normalize_dp the offsets are divided by the largest offset norm of their group (+1e-8) for 'knn', by
    `radius` otherwise — including for the spelling 'knnquery', as in the reference."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    if new_xyz is None:
        new_xyz = xyz
    if idx is not None:
        # the reference's whole body, `return` included, sits under `if idx is None` (pointops.py:125-152):
        # with a caller-provided idx it falls off the end and returns None — and so does this
        return None
    if nsample is None:                                         # "group everything": no search at all
        return xyz.transpose(1, 2).unsqueeze(2), (None if feat is None else feat.unsqueeze(2))
    use_knn = query_method in ('knn', 'knnquery')
    idx = (knnquery(nsample, xyz, new_xyz, offset, new_offset)[0] if use_knn
           else ballquery(radius, nsample, xyz, new_xyz, offset, new_offset))
    grouped_xyz = _offsets_from(xyz, new_xyz, idx)
    if normalize_dp:
        if query_method == 'knn':
            reach = grouped_xyz.norm(dim=-1, p=2, keepdim=True).max(dim=-1, keepdim=True)[0] + 1.0e-8
            grouped_xyz /= reach
        else:
            grouped_xyz /= radius
    return grouped_xyz, (None if feat is None else _gather_rows(feat, idx))


def queryandgroup(nsample, xyz, new_xyz, feat, idx, offset, new_offset, use_xyz=True):
    """kNN groups of [relative xyz | features] -> (m,nsample,3+c), or features only (pointops.py:161-184).
    `idx` (m,nsample) skips the search."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    if new_xyz is None:
        new_xyz = xyz
    if idx is None:
        idx = knnquery(nsample, xyz, new_xyz, offset, new_offset)[0]
    idx = idx.view(new_xyz.shape[0], nsample)
    grouped_feat = _gather_rows(feat, idx)
    if not use_xyz:
        return grouped_feat
    return torch.cat((_offsets_from(xyz, new_xyz, idx), grouped_feat), -1)


class Subtraction(Function):
    """input1 (n,c), input2 (n,c), idx (n,nsample) -> input1[i] - input2[idx[i,s]]  (n,nsample,c)
    (pointops.py:187-219)"""

    @staticmethod
    def forward(ctx, input1, input2, idx):
        assert input1.is_contiguous() and input2.is_contiguous()
        n, c = input1.shape
        nsample = idx.shape[-1]
        out = _buf((n, nsample, c), input1, zero=False)
        pointops_cuda.subtraction_forward_cuda(n, nsample, c, input1, input2, idx, out)
        ctx.save_for_backward(idx)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, = ctx.saved_tensors
        n, nsample, c = grad_output.shape
        g1, g2 = _buf((n, c), grad_output), _buf((n, c), grad_output)
        pointops_cuda.subtraction_backward_cuda(n, nsample, c, idx, grad_output.contiguous(), g1, g2)
        return g1, g2, None


subtraction = Subtraction.apply


class Aggregation(Function):
    """sum_s (input[idx[i,s]] + position[i,s]) * weight[i,s, ch % w_c] -> (n,c)  (pointops.py:222-256)"""

    @staticmethod
    def forward(ctx, input, position, weight, idx):
        assert input.is_contiguous() and position.is_contiguous() and weight.is_contiguous()
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        out = _buf((n, c), input)
        pointops_cuda.aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, out)
        ctx.save_for_backward(input, position, weight, idx)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, position, weight, idx = ctx.saved_tensors
        n, nsample, c = position.shape
        w_c = weight.shape[-1]
        g_in = _buf((n, c), input)
        g_pos = _buf((n, nsample, c), input, zero=False)       # written, not accumulated
        g_w = _buf((n, nsample, w_c), input)
        pointops_cuda.aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx,
                                                grad_output.contiguous(), g_in, g_pos, g_w)
        return g_in, g_pos, g_w, None


aggregation = Aggregation.apply


def _idw(xyz, new_xyz, offset, new_offset, k):
    """k nearest known points of every new point and their normalised inverse-distance weights
    w = (1/(d+1e-8)) / sum_k (1/(d+1e-8))  -> (idx (n,k) i32, w (n,k) f32)"""
    idx, dist = knnquery(k, xyz, new_xyz, offset, new_offset)
    inv = 1.0 / (dist + 1e-8)
    return idx, inv / torch.sum(inv, dim=1, keepdim=True)


def interpolation(xyz, new_xyz, feat, offset, new_offset, k=3):
    """Inverse-distance interpolation of feat (m,c) at xyz (m,3) onto new_xyz (n,3), composed from torch
    ops so that autograd differentiates it (pointops.py:259-274): one weighted row gather per neighbour
    rank, accumulated in rank order."""
    assert xyz.is_contiguous() and new_xyz.is_contiguous() and feat.is_contiguous()
    idx, weight = _idw(xyz, new_xyz, offset, new_offset, k)
    out = _buf((new_xyz.shape[0], feat.shape[1]), feat)
    for rank in range(k):
        out += feat[idx[:, rank].long(), :] * weight[:, rank].unsqueeze(-1)
    return out


class Interpolation(Function):
    """The same interpolation through the fused kernels (pointops.py:277-311)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, input, offset, new_offset, k=3):
        assert xyz.is_contiguous() and new_xyz.is_contiguous() and input.is_contiguous()
        idx, weight = _idw(xyz, new_xyz, offset, new_offset, k)
        n, c, m = new_xyz.shape[0], input.shape[1], input.shape[0]
        out = _buf((n, c), input)
        pointops_cuda.interpolation_forward_cuda(n, c, k, input, idx, weight, out)
        ctx.m, ctx.k = m, k
        ctx.save_for_backward(idx, weight)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        idx, weight = ctx.saved_tensors
        n, c = grad_output.shape
        g_in = _buf((ctx.m, c), grad_output)
        pointops_cuda.interpolation_backward_cuda(n, c, ctx.k, grad_output.contiguous(), idx, weight, g_in)
        return None, None, g_in, None, None, None


interpolation2 = Interpolation.apply
