"""Drop-in for the reference extension module ``pointops_cuda`` (Tier 1).

ref: openpoints/cpp/pointops/src/pointops_api.cpp:13-25 — every export of that table is here, with the
same argument order.  ``knnquery_cuda`` is the only one with a caller in AMContrast3D (six call sites,
all in openpoints/AMContrast3D/); the others (grouping, furthestsampling, ballquery, interpolation,
subtraction, aggregation) serve the Point-Transformer style models of OpenPoints and complete the
module surface.  Outputs are written in place, as the reference does.
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import ptr, stream


def _off(t):
    """Segment offsets as the kernels read them: contiguous int32 on the device (the reference's kernels
    reinterpret whatever they are given as int*; an int64 tensor would silently be misread)."""
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """ref: knnquery_cuda.cpp:7.  idx (m,nsample) i32 and dist2 (m,nsample) f32 (SQUARED) are
    written in place."""
    offset, new_offset = _off(offset), _off(new_offset)
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


def furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx):
    """ref: sampling_cuda.cpp furthestsampling_cuda.  tmp (n) = 1e10, idx (new_offset[-1]) i32 receives
    global indices.  The offsets are read on the host (the reference's wrapper does the same to find
    n_max, pointops.py:20-23): equal-sized consecutive segments then go out as one cluster launch."""
    h_off = offset.detach().to("cpu", torch.int32).contiguous()
    h_new = new_offset.detach().to("cpu", torch.int32).contiguous()
    with _capi.guard(xyz):
        _capi.call("amc3d_pointops_furthestsampling", int(b), int(n_max), ptr(xyz), h_off.data_ptr(),
                   h_new.data_ptr(), ptr(tmp), ptr(idx), stream(xyz))


def ballquery_cuda(m, radius, nsample, xyz, new_xyz, offset, new_offset, idx):
    """ref: ballquery_cuda.cpp:35.  idx (m,nsample) i32, zero-filled by the caller."""
    offset, new_offset = _off(offset), _off(new_offset)
    with _capi.guard(xyz):
        _capi.call("amc3d_pointops_ballquery", int(xyz.shape[0]), int(m), int(offset.shape[0]), float(radius),
                   int(nsample), ptr(xyz), ptr(new_xyz), ptr(offset), ptr(new_offset), ptr(idx), stream(xyz))
    return 1


def interpolation_forward_cuda(n, c, k, input, idx, weight, output):
    """ref: interpolation_cuda.cpp interpolation_forward_cuda (output accumulates)"""
    with _capi.guard(input):
        _capi.call("amc3d_pointops_interpolation_forward", int(n), int(c), int(k), ptr(input), ptr(idx),
                   ptr(weight), ptr(output), stream(input))


def interpolation_backward_cuda(n, c, k, grad_output, idx, weight, grad_input):
    """ref: interpolation_cuda.cpp interpolation_backward_cuda (grad_input pre-zeroed)"""
    with _capi.guard(grad_output):
        _capi.call("amc3d_pointops_interpolation_backward", int(n), int(c), int(k), ptr(grad_output), ptr(idx),
                   ptr(weight), ptr(grad_input), stream(grad_output))


def subtraction_forward_cuda(n, nsample, c, input1, input2, idx, output):
    """ref: subtraction_cuda.cpp subtraction_forward_cuda"""
    with _capi.guard(input1):
        _capi.call("amc3d_pointops_subtraction_forward", int(n), int(nsample), int(c), ptr(input1), ptr(input2),
                   ptr(idx), ptr(output), stream(input1))


def subtraction_backward_cuda(n, nsample, c, idx, grad_output, grad_input1, grad_input2):
    """ref: subtraction_cuda.cpp subtraction_backward_cuda (both gradients pre-zeroed)"""
    with _capi.guard(grad_output):
        _capi.call("amc3d_pointops_subtraction_backward", int(n), int(nsample), int(c), ptr(idx),
                   ptr(grad_output), ptr(grad_input1), ptr(grad_input2), stream(grad_output))


def aggregation_forward_cuda(n, nsample, c, w_c, input, position, weight, idx, output):
    """ref: aggregation_cuda.cpp aggregation_forward_cuda (output accumulates)"""
    with _capi.guard(input):
        _capi.call("amc3d_pointops_aggregation_forward", int(n), int(nsample), int(c), int(w_c), ptr(input),
                   ptr(position), ptr(weight), ptr(idx), ptr(output), stream(input))


def aggregation_backward_cuda(n, nsample, c, w_c, input, position, weight, idx, grad_output, grad_input,
                              grad_position, grad_weight):
    """ref: aggregation_cuda.cpp aggregation_backward_cuda (grad_input, grad_weight pre-zeroed)"""
    with _capi.guard(input):
        _capi.call("amc3d_pointops_aggregation_backward", int(n), int(nsample), int(c), int(w_c), ptr(input),
                   ptr(position), ptr(weight), ptr(idx), ptr(grad_output), ptr(grad_input), ptr(grad_position),
                   ptr(grad_weight), stream(input))
