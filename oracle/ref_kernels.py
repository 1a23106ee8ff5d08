"""ctypes binding of oracle/_ref/libref_kernels.so: the REFERENCE's own CUDA kernels.

TEST INFRASTRUCTURE ONLY.  The library is compiled by `make -C oracle ref` from the unmodified
sources under /root/reference (never copied here) and travels to the GPU box as a built file.
It is used (a) to pin ops_oracle.c against the real reference on a B200 and (b) as the
"reference CUDA recompiled for sm_100" GPU baseline in bench.py's extra report.  The reference
launchers use the legacy default stream; callers must synchronise around them.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_kernels.so")
LIB_POP = os.path.join(HERE, "_ref", "libref_pointops.so")     # the remaining pointops kernels
REF_ROOT = "/root/reference"
_lib = None
_lib_pop = None


def available() -> bool:
    return os.path.exists(LIB)


def build() -> str | None:
    """Build from /root/reference when it exists (the container); the GPU box uses the prebuilt file."""
    if os.path.isdir(REF_ROOT) and not (os.path.exists(LIB) and os.path.exists(LIB_POP)):
        subprocess.run(["make", "-C", HERE, "-j6", "ref"], check=True, capture_output=True)
    return LIB if os.path.exists(LIB) else None


def _load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_kernels.so is missing (make -C oracle ref)")
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(t):
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def sync():
    rc = _load().ref_sync()
    if rc != 0:
        raise RuntimeError(f"reference kernel failed: cudaError {rc}")


def fps(xyz, npoint):
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    idx = torch.zeros((B, npoint), dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _load().ref_fps(B, N, int(npoint), _p(xyz), _p(temp), _p(idx))
    sync()
    return idx, temp


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _load().ref_ball_query(B, N, M, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx))
    sync()
    return idx


def group_points(points, idx):
    B, C, N = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((B, C, npoints, nsample), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _load().ref_group_points(B, C, N, npoints, nsample, _p(points), _p(idx), _p(out))
    sync()
    return out


def group_points_grad(grad_out, idx, N):
    B, C, npoints, nsample = grad_out.shape
    gp = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    _load().ref_group_points_grad(B, C, int(N), npoints, nsample, _p(grad_out), _p(idx), _p(gp))
    sync()
    return gp


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.zeros((B, n, 3), dtype=torch.int32, device=unknown.device)
    torch.cuda.synchronize()
    _load().ref_three_nn(B, n, m, _p(unknown), _p(known), _p(dist2), _p(idx))
    sync()
    return dist2, idx


def three_interpolate(points, idx, weight):
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _load().ref_three_interpolate(B, C, m, n, _p(points), _p(idx), _p(weight), _p(out))
    sync()
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, n = grad_out.shape
    gp = torch.zeros((B, C, m), dtype=torch.float32, device=grad_out.device)
    torch.cuda.synchronize()
    _load().ref_three_interpolate_grad(B, C, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(gp))
    sync()
    return gp


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    _load().ref_knnquery(m, int(nsample), _p(xyz), _p(new_xyz), _p(offset), _p(new_offset), _p(idx), _p(dist2))
    sync()
    return idx, dist2


# ---- the remaining pointops kernels (oracle/_ref/libref_pointops.so) -----------------------------------

def pointops_available() -> bool:
    return os.path.exists(LIB_POP)


def _pop():
    global _lib_pop
    if _lib_pop is None:
        if not pointops_available():
            raise RuntimeError("oracle/_ref/libref_pointops.so is missing (make -C oracle ref)")
        _lib_pop = ctypes.CDLL(LIB_POP)
    return _lib_pop


def _pop_run(name, *args):
    torch.cuda.synchronize()
    getattr(_pop(), name)(*args)
    rc = _pop().ref_pop_sync()
    if rc != 0:
        raise RuntimeError(f"reference kernel {name} failed: cudaError {rc}")


def pop_furthestsampling(xyz, offset, new_offset):
    """pointops.py:10-28: n_max over the segments, tmp = 1e10"""
    h = offset.cpu().long()
    n_max = int(torch.diff(h, prepend=h.new_zeros(1)).max())
    idx = torch.zeros((int(new_offset[-1]),), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    _pop_run("ref_pop_fps", int(offset.shape[0]), n_max, _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))
    return idx, tmp


def pop_ballquery(radius, nsample, xyz, new_xyz, offset, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    _pop_run("ref_pop_ballquery", m, ctypes.c_float(radius), int(nsample), _p(xyz), _p(new_xyz), _p(offset),
             _p(new_offset), _p(idx))
    return idx


def pop_interpolation_fwd(input, idx, weight):
    n, k = idx.shape
    c = input.shape[1]
    out = torch.zeros((n, c), dtype=torch.float32, device=input.device)
    _pop_run("ref_pop_interpolation_fwd", n, c, k, _p(input), _p(idx), _p(weight), _p(out))
    return out


def pop_interpolation_bwd(grad_output, idx, weight, m):
    n, k = idx.shape
    c = grad_output.shape[1]
    g = torch.zeros((int(m), c), dtype=torch.float32, device=grad_output.device)
    _pop_run("ref_pop_interpolation_bwd", n, c, k, _p(grad_output), _p(idx), _p(weight), _p(g))
    return g


def pop_subtraction_fwd(input1, input2, idx):
    n, c = input1.shape
    ns = idx.shape[1]
    out = torch.zeros((n, ns, c), dtype=torch.float32, device=input1.device)
    _pop_run("ref_pop_subtraction_fwd", n, ns, c, _p(input1), _p(input2), _p(idx), _p(out))
    return out


def pop_subtraction_bwd(idx, grad_output):
    n, ns, c = grad_output.shape
    g1 = torch.zeros((n, c), dtype=torch.float32, device=grad_output.device)
    g2 = torch.zeros((n, c), dtype=torch.float32, device=grad_output.device)
    _pop_run("ref_pop_subtraction_bwd", n, ns, c, _p(idx), _p(grad_output), _p(g1), _p(g2))
    return g1, g2


def pop_aggregation_fwd(input, position, weight, idx):
    n, ns, c = position.shape
    w_c = weight.shape[-1]
    out = torch.zeros((n, c), dtype=torch.float32, device=input.device)
    _pop_run("ref_pop_aggregation_fwd", n, ns, c, w_c, _p(input), _p(position), _p(weight), _p(idx), _p(out))
    return out


def pop_aggregation_bwd(input, position, weight, idx, grad_output):
    n, ns, c = position.shape
    w_c = weight.shape[-1]
    gi, gp, gw = torch.zeros_like(input), torch.zeros_like(position), torch.zeros_like(weight)
    _pop_run("ref_pop_aggregation_bwd", n, ns, c, w_c, _p(input), _p(position), _p(weight), _p(idx), _p(grad_output),
             _p(gi), _p(gp), _p(gw))
    return gi, gp, gw
