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
REF_ROOT = "/root/reference"
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def build() -> str | None:
    """Build from /root/reference when it exists (the container); the GPU box uses the prebuilt file."""
    if os.path.isdir(REF_ROOT) and not os.path.exists(LIB):
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
