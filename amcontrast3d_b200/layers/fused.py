"""The fused  grouping -> 1x1 conv -> BatchNorm -> ReLU -> max  operator behind PointNeXt's LocalAggregation /
SetAbstraction modules.

ref: openpoints/models/backbone/pointnext_AA.py:57-63 (LocalAggregation.forward: grouper, 'dp_fj' concatenation,
convs, max-pool) and :139-170 (SetAbstraction.forward: FPS, gather, the same four steps);
openpoints/models/layers/group.py:235-255 (QueryAndGroup), conv.py:24-61 (conv-norm-act block).

`fused_group_conv_bn_relu_max` is the tensor-level operator; `local_aggregation_forward` /
`set_abstraction_forward` take a module built by the reference's own constructors (or any module with the same
attributes: .grouper, .convs = Sequential(Sequential(Conv2d 1x1 no bias, BatchNorm2d, ReLU)), ...) and run its
forward through the operator — same parameters, same running statistics, same outputs; `compat.install(tier=4)`
binds them as the forward of the reference's classes.  The (B, 3+C, M, nsample) grouped tensor does not exist at
any point of the forward or the backward.

Precision: the reference's Conv2d runs through cuDNN, which by torch's default (`torch.backends.cudnn.allow_tf32
= True`) reads its operands as TF32 on Ampere and later.  precision='tf32' does the same on the tcgen05 tensor
cores; precision='tf32x3' is the error-compensated three-product scheme, FP32-faithful to ~1e-6, and is what the
parity tests against the CPU-generated golden vectors use.  The default follows torch's switch.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _capi
from .._capi import ptr, stream
from .group import ball_query
from .subsample import furthest_point_sample

_PREC = {"tf32": 1, "tf32x3": 3}


def default_precision() -> str:
    return "tf32" if torch.backends.cudnn.allow_tf32 else "tf32x3"


def supported(c_in: int, nsample: int) -> bool:
    """shapes the tensor-core kernels are built for"""
    return c_in % 8 == 0 and c_in >= 8 and nsample in (16, 32)


def pack_weight(weight: torch.Tensor, c: int) -> torch.Tensor:
    """conv weight (O, 3 + C[, 1, 1]), input channels ordered [dp | features] as torch.cat((dp, fj), 1) feeds them
    -> (O, C + 8) = [features | dp | 0 0 0 0 0]: the feature columns first, so that a neighbour's row of the
    channel-contiguous feature copy lands 16-byte aligned in the operand tile."""
    w = weight.reshape(weight.shape[0], -1)
    assert w.shape[1] == c + 3, (w.shape, c)
    return torch.cat([w[:, 3:], w[:, :3], w.new_zeros(w.shape[0], 5)], 1).contiguous()


def transpose_bcn(f: torch.Tensor) -> torch.Tensor:
    """(B, C, N) -> (B, N, C) contiguous"""
    B, C, N = f.shape
    out = torch.empty((B, N, C), dtype=torch.float32, device=f.device)
    with _capi.guard(f):
        _capi.call("amc3d_transpose_batched", B, C, N, ptr(f), ptr(out), stream(f))
    return out


class FusedGroupConvBNReLUMax(Function):
    """out (B,O,M), batch mean (O), biased batch variance (O)  <-  features (B,C,N), conv weight (O,3+C), gamma, beta"""

    @staticmethod
    def forward(ctx, features, weight, gamma, beta, query_xyz, support_xyz, idx, radius, normalize_dp, eps, precision):
        assert features.is_cuda and features.dtype == torch.float32
        features = features.contiguous()
        B, C, N = features.shape
        M, ns = idx.shape[1], idx.shape[2]
        O = weight.shape[0]
        dev = features.device
        fT = transpose_bcn(features)
        wp = pack_weight(weight.detach().float(), C)
        tiles = ((O + 127) // 128) * ((C + 8 + 31) // 32) * 4096 * (2 if precision == "tf32x3" else 1)
        wt = torch.empty((tiles,), dtype=torch.float32, device=dev)
        gamma_c, beta_c = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        ysel = torch.empty((B * M, O), dtype=torch.float32, device=dev)
        arg = torch.empty((B * M, O), dtype=torch.uint8, device=dev)
        sums = torch.empty((128 * O,), dtype=torch.float64, device=dev)
        mean = torch.empty((O,), dtype=torch.float32, device=dev)
        var = torch.empty_like(mean)
        invstd = torch.empty_like(mean)
        out = torch.empty((B, O, M), dtype=torch.float32, device=dev)
        with _capi.guard(features):
            _capi.call("amc3d_fused_sa_forward", B, N, M, C, O, ns, float(radius), int(bool(normalize_dp)),
                       _PREC[precision], float(eps), ptr(fT), ptr(support_xyz), ptr(query_xyz), ptr(idx), ptr(wp),
                       ptr(wt), ptr(gamma_c), ptr(beta_c), ptr(ysel), ptr(arg), ptr(sums), ptr(mean), ptr(var), ptr(invstd),
                       ptr(out), stream(features))
        ctx.save_for_backward(fT, wp, gamma_c, query_xyz, support_xyz, idx, ysel, arg, mean, invstd, out)
        ctx.cfg = (float(radius), bool(normalize_dp), precision, weight.shape)
        ctx.mark_non_differentiable(mean, var)
        return out, mean, var

    @staticmethod
    def backward(ctx, grad_out, _gm, _gv):
        from . import _fused_backward
        return _fused_backward.backward(ctx, grad_out)


def fused_group_conv_bn_relu_max(query_xyz, support_xyz, features, idx, weight, bn: nn.BatchNorm2d, radius,
                                 normalize_dp=True, precision=None):
    """relu(bn(conv1x1(cat(dp, fj)))).max(-1)  with  (dp, fj) = QueryAndGroup(radius, ns, normalize_dp)(query, support,
    features)  and  idx = ball_query(radius, ns, support, query)  ->  (B, O, M).

    `bn` is the module's BatchNorm2d: in training mode its batch statistics are used and its running statistics
    updated exactly as nn.BatchNorm2d does (momentum, unbiased variance, num_batches_tracked)."""
    if not bn.training:
        raise NotImplementedError("the fused operator implements the training-mode step (batch statistics); "
                                  "in eval mode fold the running statistics into the conv and call the module")
    precision = precision or default_precision()
    out, mean, var = FusedGroupConvBNReLUMax.apply(features, weight, bn.weight, bn.bias, query_xyz.contiguous(),
                                                   support_xyz.contiguous(), idx.contiguous(), radius, normalize_dp,
                                                   bn.eps, precision)
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            n = idx.numel()
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var * (n / max(n - 1, 1)), alpha=mom)
    return out


# ----------------------------------------------------------------------------------------------------------------
# module-level entry points: the forward of LocalAggregation / SetAbstraction
# ----------------------------------------------------------------------------------------------------------------
def _single_block(convs):
    """(conv, bn) if `convs` is Sequential(Sequential(Conv2d 1x1 bias-free, BatchNorm2d, ReLU)), else None"""
    if len(convs) != 1:
        return None
    blk = list(convs[0])
    if len(blk) != 3:
        return None
    conv, bn, act = blk
    if not (isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d) and isinstance(act, nn.ReLU)):
        return None
    if conv.kernel_size != (1, 1) or conv.bias is not None or conv.groups != 1:
        return None
    return conv, bn


def _fusable(module, f):
    g = module.grouper
    ok = (getattr(module, "feature_type", None) == "dp_fj" and hasattr(g, "radius") and g.radius is not None
          and getattr(g, "relative_xyz", True) and not getattr(g, "normalize_by_std", False)
          and not getattr(g, "normalize_by_allstd", False) and not getattr(g, "normalize_by_allstd2", False)
          and f.is_cuda and f.dtype == torch.float32 and module.training)
    if not ok:
        return None
    cb = _single_block(module.convs)
    if cb is None or not supported(f.shape[1], int(g.nsample)) or not cb[1].training:
        return None
    return cb


def local_aggregation_forward(module, pf, precision=None):
    """LocalAggregation.forward(pf) (pointnext_AA.py:57-63) through the fused operator.  Returns None for
    configurations the operator does not cover (several conv layers, other feature types or reductions, eval
    mode, channel counts off the 8-grid): the caller runs the module's own composition."""
    p, f = pf
    cb = _fusable(module, f) if getattr(module, "reduction", "max") == "max" else None
    if cb is None:
        return None
    conv, bn = cb
    g = module.grouper
    idx = ball_query(g.radius, g.nsample, p, p)
    return fused_group_conv_bn_relu_max(p, p, f, idx, conv.weight, bn, g.radius, g.normalize_dp, precision)


def set_abstraction_forward(module, pf, precision=None):
    """SetAbstraction.forward(pf) (pointnext_AA.py:139-170) for the non-head, non-residual, strided layer:
    FPS -> gather queries -> fused operator.  -> (new_p, f), or None when not covered."""
    p, f = pf
    if module.is_head or module.all_aggr or module.use_res:
        return None
    cb = _fusable(module, f)
    if cb is None:
        return None
    conv, bn = cb
    idx_fps = furthest_point_sample(p, p.shape[1] // module.stride).long()
    new_p = torch.gather(p, 1, idx_fps.unsqueeze(-1).expand(-1, -1, 3))
    g = module.grouper
    idx = ball_query(g.radius, g.nsample, p, new_p.contiguous())
    return new_p, fused_group_conv_bn_relu_max(new_p, p, f, idx, conv.weight, bn, g.radius, g.normalize_dp, precision)


def bind(cls, which: str):
    """Route `cls.forward` (the reference's LocalAggregation / SetAbstraction class, or a subclass) through the fused
    operator, keeping the original forward for everything the operator does not cover.  Idempotent."""
    if getattr(cls, "_amc3d_fused", False):
        return cls
    orig = cls.forward
    fn = local_aggregation_forward if which == "la" else set_abstraction_forward

    def forward(self, pf):
        out = fn(self, pf)
        return orig(self, pf) if out is None else out

    forward.__doc__ = orig.__doc__
    cls._amc3d_orig_forward = orig
    cls.forward = forward
    cls._amc3d_fused = True
    return cls
