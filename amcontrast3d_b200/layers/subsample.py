"""Sub-sampling operators with the reference signatures (Tier 2).

Mirrors openpoints/models/layers/subsample.py:76-157: ``furthest_point_sample``,
``gather_operation``, ``random_sample``, ``fps``.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import pointnet2_batch_cuda as pointnet2_cuda


class FurthestPointSampling(Function):
    """ref: subsample.py:76-103.  xyz (B,N,3) contiguous f32 -> idx (B,npoint) i32; no grad."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = int(npoint)
        output = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        pointnet2_cuda.furthest_point_sampling_wrapper(B, N, npoint, xyz, temp, output)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """ref: subsample.py:109-143.  features (B,C,N), idx (B,npoint) -> (B,C,npoint)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        pointnet2_cuda.gather_points_wrapper(B, C, N, npoint, features, idx, output)
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        pointnet2_cuda.gather_points_grad_wrapper(B, C, N, npoint, grad_out.contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


def random_sample(xyz, npoint):
    """ref: subsample.py:70-73"""
    B, N, _ = xyz.shape
    return torch.randint(0, N, (B, npoint), device=xyz.device)


def fps(data, number):
    """ref: subsample.py:149-157.  data (B,N,C>=3) -> the `number` FPS-selected rows."""
    fps_idx = furthest_point_sample(data[:, :, :3].contiguous(), number)
    return torch.gather(data, 1, fps_idx.unsqueeze(-1).long().expand(-1, -1, data.shape[-1]))
