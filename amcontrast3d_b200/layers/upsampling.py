"""three_nn / three_interpolate with the reference signatures (Tier 2).

Mirrors openpoints/models/layers/upsampling.py:11-40 (ThreeNN), :43-89 (ThreeInterpolate),
:92-102 (three_interpolation).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch.autograd import Function

from .. import pointnet2_batch_cuda as pointnet2_cuda


class ThreeNN(Function):
    """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) = sqrt of the squared distances, idx i32)"""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
        pointnet2_cuda.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,C,m), idx (B,n,3), weight (B,n,3) -> (B,C,n)"""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        output = torch.empty((B, c, n), dtype=torch.float32, device=features.device)
        pointnet2_cuda.three_interpolate_wrapper(B, c, m, n, features, idx, weight, output)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.empty((B, c, m), dtype=torch.float32, device=grad_out.device)
        pointnet2_cuda.three_interpolate_grad_set(B, c, n, m, grad_out.contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


def three_interpolation(unknown_xyz, known_xyz, know_feat, nn=None):
    """Inverse-distance interpolation from the 3 nearest known points (ref: upsampling.py:92-102).
    unknown_xyz (B,n,3), known_xyz (B,m,3), know_feat (B,C,m) -> (B,C,n).
    `nn` (not in the reference signature, optional): a precomputed three_nn(unknown_xyz, known_xyz)."""
    dist, idx = three_nn(unknown_xyz, known_xyz) if nn is None else nn
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm
    return three_interpolate(know_feat, idx, weight)
