"""Grouping operators and grouper modules with the reference signatures (Tier 2).

Mirrors openpoints/models/layers/group.py:76-117 (GroupingOperation), :177-203 (BallQuery),
:206-255 (QueryAndGroup), :258-272 (GroupAll), :338-352 (create_grouper).  ``KNNGroup``
(:275-322) depends on the reference's torch cdist+topk ``KNN`` layer, which no shipped
config selects (group_args.NAME is 'ballquery'); it is reproduced on top of
``pointops.knnquery``-free torch ops only as far as the signature goes.
"""
from __future__ import annotations

import copy
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import pointnet2_batch_cuda as pointnet2_cuda


class GroupingOperation(Function):
    """features (B,C,N), idx (B,npoint,nsample) i32 -> (B,C,npoint,nsample); AMP inputs are
    cast to fp32 as in the reference (group.py:79)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, nfeatures, nsample), dtype=torch.float32, device=features.device)
        pointnet2_cuda.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, output)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor) -> Tuple[torch.Tensor, None]:
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        # the reference zero-fills and accumulates (group.py:111-113); writing the result is the same thing
        grad_features = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
        pointnet2_cuda.group_points_grad_set(B, C, N, npoint, nsample, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


def torch_grouping_operation(features, idx):
    """Pure-torch equivalent kept for API parity (ref: group.py:120-137)."""
    B, C = features.shape[:2]
    flat = idx.reshape(B, 1, -1).expand(-1, C, -1).long()
    return features.gather(2, flat).reshape(B, C, idx.shape[1], idx.shape[2])


class BallQuery(Function):
    """radius, nsample, xyz (B,N,3) support, new_xyz (B,npoint,3) centres -> idx
    (B,npoint,nsample) i32, zero-initialised like the reference (group.py:194)."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros((B, npoint, int(nsample)), dtype=torch.int32, device=xyz.device)
        pointnet2_cuda.ball_query_wrapper(B, N, npoint, radius, int(nsample), new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def _group_xyz_relative(support_xyz, query_xyz, idx, subtract, radius):
    import numpy as np

    from .. import _capi
    B, N, _ = support_xyz.shape
    M, ns = idx.shape[1], idx.shape[2]
    out = torch.empty((B, 3, M, ns), dtype=torch.float32, device=support_xyz.device)
    inv = float(np.float32(1.0) / np.float32(radius)) if radius is not None else 0.0
    with _capi.guard(support_xyz):
        _capi.call("amc3d_group_xyz_relative", B, N, M, ns, int(bool(subtract)), inv, _capi.ptr(support_xyz),
                   _capi.ptr(query_xyz), _capi.ptr(idx), _capi.ptr(out), _capi.stream(support_xyz))
    return out


class QueryAndGroup(nn.Module):
    """ref: group.py:206-255.  forward(query_xyz (B,M,3), support_xyz (B,N,3), features (B,C,N))
    -> (grouped relative xyz (B,3,M,ns), grouped features (B,C,M,ns))."""

    def __init__(self, radius: float, nsample: int, relative_xyz=True, normalize_dp=False,
                 normalize_by_std=False, normalize_by_allstd=False, normalize_by_allstd2=False,
                 return_only_idx=False, **kwargs):
        super().__init__()
        self.radius, self.nsample = radius, nsample
        self.normalize_dp = normalize_dp
        self.normalize_by_std = normalize_by_std
        self.normalize_by_allstd = normalize_by_allstd
        self.normalize_by_allstd2 = normalize_by_allstd2
        assert self.normalize_dp + self.normalize_by_std + self.normalize_by_allstd < 2
        self.relative_xyz = relative_xyz
        self.return_only_idx = return_only_idx

    def forward(self, query_xyz, support_xyz, features=None, idx=None):
        """`idx` (not in the reference signature, optional): the result of
        ball_query(self.radius, self.nsample, support_xyz, query_xyz) when the caller has already computed
        it, e.g. on a geometry stream running ahead of the feature path (replay.py)."""
        if idx is None:
            idx = ball_query(self.radius, self.nsample, support_xyz, query_xyz)
        if self.return_only_idx:
            return idx
        plain = not (self.normalize_by_std or self.normalize_by_allstd or self.normalize_by_allstd2)
        if (plain and support_xyz.is_cuda and not support_xyz.requires_grad and not query_xyz.requires_grad
                and support_xyz.dtype == torch.float32 and query_xyz.dtype == torch.float32
                and support_xyz.is_contiguous() and query_xyz.is_contiguous() and idx.is_contiguous()):
            # coordinates carry no gradient here (they never do in PointNeXt): the transpose, the C = 3
            # grouping, the subtraction and the division of the reference collapse into one kernel with
            # bit-identical results (torch divides by a Python scalar as a * (1/r) on CUDA)
            grouped_xyz = _group_xyz_relative(support_xyz, query_xyz, idx, self.relative_xyz,
                                              self.radius if (self.relative_xyz and self.normalize_dp) else None)
        else:
            grouped_xyz = grouping_operation(support_xyz.transpose(1, 2).contiguous(), idx)
            if self.relative_xyz:
                grouped_xyz = grouped_xyz - query_xyz.transpose(1, 2).unsqueeze(-1)
                if self.normalize_dp:
                    grouped_xyz /= self.radius
        grouped_features = grouping_operation(features, idx) if features is not None else None
        return grouped_xyz, grouped_features


class GroupAll(nn.Module):
    """ref: group.py:258-272"""

    def forward(self, new_xyz, xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        grouped_features = features.unsqueeze(2) if features is not None else None
        return grouped_xyz, grouped_features


class KNNGroup(nn.Module):
    """ref: group.py:275-322.  The kNN itself is torch cdist+topk in the reference (knn.py:7-20),
    not an extension op, and is not selected by any shipped config."""

    def __init__(self, nsample: int, relative_xyz=True, normalize_dp=False, return_only_idx=False, **kwargs):
        super().__init__()
        self.nsample = nsample
        self.relative_xyz = relative_xyz
        self.normalize_dp = normalize_dp
        self.return_only_idx = return_only_idx

    def forward(self, query_xyz, support_xyz, features=None):
        idx = torch.cdist(query_xyz, support_xyz).topk(self.nsample, dim=-1, largest=False)[1]
        if self.return_only_idx:
            return idx
        idx = idx.int().contiguous()
        grouped_xyz = grouping_operation(support_xyz.transpose(1, 2).contiguous(), idx)
        if self.relative_xyz:
            grouped_xyz = grouped_xyz - query_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_dp:
            grouped_xyz = grouped_xyz / torch.amax(torch.sqrt(torch.sum(grouped_xyz ** 2, dim=1)),
                                                   dim=(1, 2)).view(-1, 1, 1, 1)
        if features is not None:
            return grouped_xyz, grouping_operation(features, idx)
        return grouped_xyz, None


def create_grouper(group_args):
    """ref: group.py:338-352"""
    args = copy.deepcopy(dict(group_args))
    method = args.pop("NAME", "ballquery")
    radius = args.pop("radius", 0.1)
    nsample = args.pop("nsample", 20)
    if nsample is None:
        return GroupAll()
    if method == "ballquery":
        return QueryAndGroup(radius, nsample, **args)
    if method == "knn":
        return KNNGroup(nsample, **args)
    raise ValueError(f"unknown grouper {method!r}")
