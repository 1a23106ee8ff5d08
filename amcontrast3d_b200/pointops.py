"""Mirror of openpoints/cpp/pointops/functions/pointops.py for the exports on the hot path.

``knnquery(nsample, xyz, new_xyz, offset, new_offset) -> (idx i32 (m,nsample), dist f32)``
(pointops.py:32-56) is what AMContrast3D calls; ``grouping`` (pointops.py, Grouping) is
provided on top of the packed (n,c) gather kernel.
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
