"""Drop-in for the reference extension module ``pointnet2_batch_cuda`` (Tier 1).

Same function names, argument order and in-place output convention as the pybind table in
openpoints/cpp/pointnet2_batch/src/pointnet2_api.cpp:10-24: the caller allocates every
output (and pre-zeroes gradient buffers / ball-query indices), the functions borrow the raw
pointers for one launch.  Each function forwards to the C-ABI of include/amc3d.h on the
tensors' device and torch's current stream.  No CPU path.
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import ptr, stream


def _guard(t):
    return _capi.guard(t)




def _workspace(like: torch.Tensor, numel: int) -> torch.Tensor:
    return torch.empty(numel, dtype=torch.float32, device=like.device)


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    """ref: ball_query.cpp:29 ball_query_wrapper_fast"""
    with _guard(xyz):
        _capi.call("amc3d_ball_query", b, n, m, float(radius), int(nsample), ptr(new_xyz), ptr(xyz),
                   ptr(idx), stream(xyz))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    """ref: group_points.cpp:25 group_points_wrapper_fast"""
    with _guard(points):
        ws = _workspace(points, b * n * c) if c >= 8 else None
        _capi.call("amc3d_group_points_ws", b, c, n, npoints, nsample, ptr(points), ptr(idx), ptr(out),
                   ptr(ws), stream(points))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    """ref: group_points.cpp:13 group_points_grad_wrapper_fast (grad_points pre-zeroed)"""
    with _guard(grad_out):
        ws = _workspace(grad_out, b * n * c) if c >= 8 else None
        _capi.call("amc3d_group_points_grad_ws", b, c, n, npoints, nsample, ptr(grad_out), ptr(idx),
                   ptr(grad_points), ptr(ws), stream(grad_out))
    return 1


def group_points_grad_set(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    """Not in the reference module: group_points_grad_wrapper for an UNINITIALISED grad_points (written, not
    accumulated) — what layers.GroupingOperation.backward uses, saving the zero-fill and one read of it."""
    with _guard(grad_out):
        ws = _workspace(grad_out, b * n * c) if c >= 8 else None
        _capi.call("amc3d_group_points_grad_ws_set", b, c, n, npoints, nsample, ptr(grad_out), ptr(idx),
                   ptr(grad_points), ptr(ws), stream(grad_out))
    return 1


def three_interpolate_grad_set(b, c, n, m, grad_out, idx, weight, grad_points):
    """Not in the reference module: three_interpolate_grad_wrapper for an uninitialised grad_points."""
    with _guard(grad_out):
        ws = _workspace(grad_out, b * m * c) if c >= 8 else None
        _capi.call("amc3d_three_interpolate_grad_ws_set", b, c, n, m, ptr(grad_out), ptr(idx), ptr(weight),
                   ptr(grad_points), ptr(ws), stream(grad_out))


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    """ref: sampling.cpp:16 gather_points_wrapper_fast"""
    with _guard(points):
        _capi.call("amc3d_gather_points", b, c, n, npoints, ptr(points), ptr(idx), ptr(out), stream(points))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    """ref: sampling.cpp:27 gather_points_grad_wrapper_fast (grad_points pre-zeroed)"""
    with _guard(grad_out):
        _capi.call("amc3d_gather_points_grad", b, c, n, npoints, ptr(grad_out), ptr(idx),
                   ptr(grad_points), stream(grad_out))
    return 1


def furthest_point_sampling_wrapper(b, n, m, xyz, temp, idx):
    """ref: sampling.cpp:39 furthest_point_sampling_wrapper (temp = 1e10 on entry, clobbered)"""
    with _guard(xyz):
        _capi.call("amc3d_furthest_point_sampling", b, n, m, ptr(xyz), ptr(temp), ptr(idx), stream(xyz))
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    """ref: interpolate.cpp:20 three_nn_wrapper_fast (dist2 is SQUARED)"""
    with _guard(unknown):
        _capi.call("amc3d_three_nn", b, n, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), stream(unknown))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    """ref: interpolate.cpp:31 three_interpolate_wrapper_fast"""
    with _guard(points):
        ws = _workspace(points, b * m * c) if c >= 8 else None
        _capi.call("amc3d_three_interpolate_ws", b, c, m, n, ptr(points), ptr(idx), ptr(weight), ptr(out),
                   ptr(ws), stream(points))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    """ref: interpolate.cpp:45 three_interpolate_grad_wrapper_fast (grad_points pre-zeroed)"""
    with _guard(grad_out):
        ws = _workspace(grad_out, b * m * c) if c >= 8 else None
        _capi.call("amc3d_three_interpolate_grad_ws", b, c, n, m, ptr(grad_out), ptr(idx), ptr(weight),
                   ptr(grad_points), ptr(ws), stream(grad_out))
