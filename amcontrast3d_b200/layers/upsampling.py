"""Feature propagation operators of the decoder: three_nn, three_interpolate, three_interpolation.

Public names and call signatures are those of openpoints/models/layers/upsampling.py (ThreeNN :11-40,
ThreeInterpolate :43-89, three_interpolation :92-102) so that PointNeXt's FeaturePropogation calls them
unchanged; what runs underneath is libamc3d (culled exact 3-NN search, TMA-staged interpolation).
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import pointnet2_batch_cuda as _ext

_F32 = torch.float32


def _new(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _require_contiguous(**tensors):
    for name, t in tensors.items():
        assert t.is_contiguous(), f"{name} must be contiguous"


class ThreeNN(Function):
    """For every `unknown` point the three nearest `known` points.

    unknown (B,n,3) f32, known (B,m,3) f32  ->  (dist (B,n,3) f32, idx (B,n,3) i32), dist being the
    Euclidean distance (the kernel returns squared distances; the square root is taken here, as the
    reference's wrapper does).  Neither output is differentiable."""

    @staticmethod
    def forward(ctx, unknown, known):
        _require_contiguous(unknown=unknown, known=known)
        batch, n_unknown = unknown.shape[0], unknown.shape[1]
        n_known = known.shape[1]
        sq = _new((batch, n_unknown, 3), _F32, unknown)
        nearest = _new((batch, n_unknown, 3), torch.int32, unknown)
        _ext.three_nn_wrapper(batch, n_unknown, n_known, unknown, known, sq, nearest)
        dist = sq.sqrt_()
        ctx.mark_non_differentiable(dist, nearest)
        return dist, nearest

    @staticmethod
    def backward(ctx, *unused):
        return None, None


class ThreeInterpolate(Function):
    """out[b,c,i] = sum_t weight[b,i,t] * features[b,c,idx[b,i,t]].

    features (B,C,m), idx (B,n,3) i32, weight (B,n,3) -> (B,C,n).  Under autocast the inputs are cast to
    FP32 first (the reference decorates its forward the same way).  The gradient flows to `features`."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=_F32)
    def forward(ctx, features, idx, weight):
        _require_contiguous(features=features, idx=idx, weight=weight)
        batch, channels, n_known = features.shape
        n_out = idx.shape[1]
        out = _new((batch, channels, n_out), _F32, features)
        _ext.three_interpolate_wrapper(batch, channels, n_known, n_out, features, idx, weight, out)
        ctx.save_for_backward(idx, weight)
        ctx.n_known = n_known
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        batch, channels, n_out = grad_out.shape
        # written, not accumulated: no zero-fill needed (the reference zero-fills and atomically adds)
        grad_features = _new((batch, channels, ctx.n_known), _F32, grad_out)
        _ext.three_interpolate_grad_set(batch, channels, n_out, ctx.n_known, grad_out.contiguous(), idx, weight,
                                        grad_features)
        return grad_features, None, None


three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply


def three_interpolation(unknown_xyz, known_xyz, know_feat, nn=None):
    """Inverse-distance-weighted interpolation of `know_feat` (B,C,m), given at `known_xyz` (B,m,3), onto
    `unknown_xyz` (B,n,3) -> (B,C,n): weights 1/(d + 1e-8) over the three nearest known points, normalised
    to sum to one.

    `nn` is an extension of the reference signature: a precomputed `three_nn(unknown_xyz, known_xyz)`,
    for callers that run the search ahead of the feature path (replay.py's geometry stream)."""
    dist, nearest = nn if nn is not None else three_nn(unknown_xyz, known_xyz)
    inv = 1.0 / (dist + 1e-8)
    weight = inv / inv.sum(dim=2, keepdim=True)
    return three_interpolate(know_feat, nearest, weight)
