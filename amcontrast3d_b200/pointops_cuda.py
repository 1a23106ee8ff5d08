"""Drop-in for the reference extension module ``pointops_cuda`` (Tier 1).

ref: openpoints/cpp/pointops/src/pointops_api.cpp:13-25.  ``knnquery_cuda`` is the only
export with a caller in AMContrast3D (six call sites, all in openpoints/AMContrast3D/);
``grouping_forward_cuda`` / ``grouping_backward_cuda`` are provided as well.  The remaining
Point-Transformer exports (furthestsampling, ballquery, interpolation, subtraction,
aggregation) have no caller in the reference and are listed as "next" in DESIGN.md — calling
them raises NotImplementedError rather than silently doing something else.
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import ptr, stream


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """ref: knnquery_cuda.cpp:7.  idx (m,nsample) i32 and dist2 (m,nsample) f32 (SQUARED) are
    written in place."""
    with _capi.guard(xyz):
        _capi.call("amc3d_knnquery", int(xyz.shape[0]), int(m), int(offset.shape[0]), int(nsample),
                   ptr(xyz), ptr(new_xyz), ptr(offset), ptr(new_offset), ptr(idx), ptr(dist2),
                   stream(xyz))


def grouping_forward_cuda(m, nsample, c, input, idx, output):
    """ref: grouping_cuda.cpp grouping_forward_cuda: output[i,s,:] = input[idx[i,s],:]"""
    with _capi.guard(input):
        _capi.call("amc3d_grouping_forward", int(m), int(nsample), int(c), ptr(input), ptr(idx),
                   ptr(output), stream(input))


def grouping_backward_cuda(m, nsample, c, grad_output, idx, grad_input):
    """ref: grouping_cuda.cpp grouping_backward_cuda (grad_input pre-zeroed)"""
    with _capi.guard(grad_output):
        _capi.call("amc3d_grouping_backward", int(m), int(nsample), int(c), ptr(grad_output), ptr(idx),
                   ptr(grad_input), stream(grad_output))


def _not_on_path(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(
            f"pointops_cuda.{name} has no caller in AMContrast3D and is not part of the B200 hot path "
            "(DESIGN.md, 'next')")
    fn.__name__ = name
    return fn


for _n in ("furthestsampling_cuda", "ballquery_cuda", "interpolation_forward_cuda",
           "interpolation_backward_cuda", "subtraction_forward_cuda", "subtraction_backward_cuda",
           "aggregation_forward_cuda", "aggregation_backward_cuda"):
    globals()[_n] = _not_on_path(_n)
