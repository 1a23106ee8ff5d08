"""CPU restatement of  QueryAndGroup -> 1x1 conv -> BatchNorm (training statistics) -> ReLU -> max over the
neighbourhood, the operator PointNeXt's LocalAggregation / SetAbstraction compose from separate modules
(ref: openpoints/models/backbone/pointnext_AA.py:57-63 and :147-166; grouper openpoints/models/layers/
group.py:235-255; feature_type 'dp_fj' group.py:324-325; conv block conv-norm-act with bias-free Conv2d).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  It is the parity target of the fused operator that is
next on the plan: pinned against the reference's own modules run on CPU
(tests/golden/fused_golden.npz, tests/golden/make_fused_golden.py) by tests/test_oracle_fused_golden.py.
Gradients come from autograd over this composition.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops_oracle as oo


def grouped_input(query_xyz, support_xyz, f, radius, nsample, normalize_dp=True):
    """-> x (B, 3 + C, npoint, nsample): [relative (optionally radius-normalised) xyz | neighbour features]
    (group.py:244-255 followed by the 'dp_fj' concatenation), and idx (B, npoint, nsample) i32."""
    q, s = query_xyz.detach().numpy(), support_xyz.detach().numpy()
    idx = torch.from_numpy(oo.ball_query(radius, nsample, s, q)).long()              # (B, npoint, nsample)
    B, npoint, ns = idx.shape
    flat = idx.reshape(B, 1, npoint * ns)
    dp = torch.gather(support_xyz.transpose(1, 2), 2, flat.expand(-1, 3, -1)).reshape(B, 3, npoint, ns)
    dp = dp - query_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_dp:
        dp = dp / radius
    fj = torch.gather(f, 2, flat.expand(-1, f.shape[1], -1)).reshape(B, f.shape[1], npoint, ns)
    return torch.cat([dp, fj], 1), idx.int()


def conv_bn_relu_max(x, w, gamma, beta, eps=1e-5):
    """x (B, Cin, npoint, nsample), w (Cout, Cin) -> (B, Cout, npoint); BatchNorm2d in training mode: per
    output channel the mean and the BIASED variance over (B, npoint, nsample).  Also returns (mean, unbiased
    variance), what the module folds into its running statistics."""
    y = torch.einsum("oc,bcps->bops", w, x)
    mean = y.mean(dim=(0, 2, 3), keepdim=True)
    var = y.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    n = y.numel() // y.shape[1]
    z = (y - mean) / torch.sqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return torch.relu(z).max(dim=-1)[0], mean.flatten(), var.flatten() * (n / max(n - 1, 1))


def local_aggregation(xyz, f, w, gamma, beta, radius, nsample, eps=1e-5):
    """LocalAggregation.forward (pointnext_AA.py:57-63): queries = support = xyz"""
    x, _ = grouped_input(xyz, xyz, f, radius, nsample)
    return conv_bn_relu_max(x, w, gamma, beta, eps)


def set_abstraction(xyz, f, w, gamma, beta, stride, radius, nsample, eps=1e-5):
    """SetAbstraction.forward without residual (pointnext_AA.py:147-166): FPS to N/stride queries first.
    -> (new_xyz, features, mean, unbiased var)"""
    idx, _ = oo.fps(xyz.detach().numpy(), xyz.shape[1] // stride)
    new_xyz = torch.gather(xyz, 1, torch.from_numpy(idx).long().unsqueeze(-1).expand(-1, -1, 3))
    x, _ = grouped_input(new_xyz, xyz, f, radius, nsample)
    return (new_xyz,) + conv_bn_relu_max(x, w, gamma, beta, eps)
