"""numpy front-end of ops_oracle.c (CPU restatement of the reference operator kernels).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Every function takes and returns numpy
arrays laid out exactly like the reference's tensors and cites the reference kernel it
restates in ops_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libops_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ops_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.oracle_opt_n_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def host_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def _parallel(total, fn, threads=None):
    """Run fn(lo, hi) over [0,total) split across host threads (ctypes drops the GIL)."""
    threads = threads or host_threads()
    if total == 0:
        return
    threads = min(threads, total)
    edges = np.linspace(0, total, threads + 1).astype(np.int64)
    if threads == 1:
        fn(0, total)
        return
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: fn(int(edges[i]), int(edges[i + 1])), range(threads)))


def opt_n_threads(n: int) -> int:
    return _load().oracle_opt_n_threads(int(n))


def fps(xyz, npoint, temp=None):
    """sampling_gpu.cu:100-216.  xyz (B,N,3) -> idx (B,npoint) i32 (and the clobbered temp)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, dtype=np.float32) if temp is None else _f32(temp).copy()
    idx = np.zeros((B, npoint), dtype=np.int32)
    lib = _load()

    def one(lo, hi):
        for b in range(lo, hi):
            lib.oracle_fps(1, N, int(npoint), _p(xyz[b]), _p(temp[b]), _p(idx[b]))

    _parallel(B, one)
    return idx, temp


def ball_query(radius, nsample, xyz, new_xyz, threads=None):
    """ball_query_gpu.cu:15-51.  xyz (B,N,3) support, new_xyz (B,M,3) -> idx (B,M,nsample)."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    lib = _load()
    _parallel(B * M, lambda lo, hi: lib.oracle_ball_query_range(
        ctypes.c_longlong(lo), ctypes.c_longlong(hi), B, N, M, ctypes.c_float(radius), int(nsample),
        _p(new_xyz), _p(xyz), _p(idx)), threads)
    return idx


def three_nn(unknown, known, threads=None):
    """interpolate_gpu.cu:16-59 -> (dist2 (B,n,3) SQUARED, idx (B,n,3))."""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    lib = _load()
    _parallel(B * n, lambda lo, hi: lib.oracle_three_nn_range(
        ctypes.c_longlong(lo), ctypes.c_longlong(hi), B, n, m, _p(unknown), _p(known), _p(dist2), _p(idx)),
        threads)
    return dist2, idx


def three_interpolate(points, idx, weight):
    """interpolate_gpu.cu:84-104.  points (B,C,m), idx/weight (B,n,3) -> (B,C,n)."""
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), dtype=np.float32)
    _load().oracle_three_interpolate(B, C, m, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """interpolate_gpu.cu:127-149 -> grad_points (B,C,m)."""
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    gp = np.zeros((B, C, m), dtype=np.float32)
    _load().oracle_three_interpolate_grad(B, C, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(gp))
    return gp


def group_points(points, idx):
    """group_points_gpu.cu:53-72.  points (B,C,N), idx (B,np,ns) -> (B,C,np,ns)."""
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, npoints, nsample = idx.shape
    out = np.zeros((B, C, npoints, nsample), dtype=np.float32)
    _load().oracle_group_points(B, C, N, npoints, nsample, _p(points), _p(idx), _p(out))
    return out


def group_points_grad(grad_out, idx, N):
    """group_points_gpu.cu:14-31 -> grad_points (B,C,N)."""
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, npoints, nsample = grad_out.shape
    gp = np.zeros((B, C, N), dtype=np.float32)
    _load().oracle_group_points_grad(B, C, int(N), npoints, nsample, _p(grad_out), _p(idx), _p(gp))
    return gp


def gather_points(points, idx):
    """sampling_gpu.cu:15-31.  points (B,C,N), idx (B,m) -> (B,C,m)."""
    return group_points(points, _i32(idx)[:, :, None])[..., 0]


def gather_points_grad(grad_out, idx, N):
    """sampling_gpu.cu:53-70"""
    return group_points_grad(_f32(grad_out)[..., None], _i32(idx)[:, :, None], N)


def knnquery(nsample, xyz, new_xyz, offset, new_offset, lex=False, threads=None):
    """knnquery_cuda_kernel.cu:65-108 -> (idx (m,nsample) i32, dist2 (m,nsample) SQUARED).
    lex=False: the literal max-heap + heap sort; lex=True: (d2, index)-lexicographic top-k."""
    xyz = _f32(xyz)
    new_xyz = xyz if new_xyz is None else _f32(new_xyz)
    offset, new_offset = _i32(offset), _i32(new_offset)
    m = new_xyz.shape[0]
    nsample = int(nsample)
    assert nsample <= 128
    idx = np.zeros((m, nsample), dtype=np.int32)
    dist2 = np.zeros((m, nsample), dtype=np.float32)
    lib = _load()
    fn = lib.oracle_knnquery_lex_range if lex else lib.oracle_knnquery_range
    _parallel(m, lambda lo, hi: fn(lo, hi, nsample, int(offset.shape[0]), _p(xyz), _p(new_xyz),
                                   _p(offset), _p(new_offset), _p(idx), _p(dist2)), threads)
    return idx, dist2


# ---- the remaining pointops kernels (packed layout) --------------------------------------------------

def pop_furthestsampling(xyz, offset, new_offset):
    """sampling_cuda_kernel.cu:13-133 with the wrapper's n_max / tmp=1e10 (pointops.py:18-26)
    -> idx (new_offset[-1]) i32, global indices."""
    xyz, offset, new_offset = _f32(xyz), _i32(offset), _i32(new_offset)
    sizes = np.diff(np.concatenate([[0], offset]))
    n_max = int(sizes.max()) if sizes.size else 0
    idx = np.zeros((int(new_offset[-1]) if new_offset.size else 0,), dtype=np.int32)
    tmp = np.full((xyz.shape[0],), 1e10, dtype=np.float32)
    _load().oracle_pop_fps(int(offset.shape[0]), n_max, _p(xyz), _p(offset), _p(new_offset), _p(tmp), _p(idx))
    return idx


def pop_ballquery(radius, nsample, xyz, new_xyz, offset, new_offset, threads=None):
    """ballquery_cuda_kernel.cu:27-76 -> idx (m,nsample) i32 (rows without a hit stay zero)"""
    xyz = _f32(xyz)
    new_xyz = xyz if new_xyz is None else _f32(new_xyz)
    offset, new_offset = _i32(offset), _i32(new_offset)
    m = new_xyz.shape[0]
    idx = np.zeros((m, int(nsample)), dtype=np.int32)
    lib = _load()
    _parallel(m, lambda lo, hi: lib.oracle_pop_ballquery_range(lo, hi, ctypes.c_float(radius), int(nsample), _p(xyz),
                                                               _p(new_xyz), _p(offset), _p(new_offset), _p(idx)),
              threads)
    return idx


def pop_interpolation_fwd(input, idx, weight, output=None):
    """interpolation_cuda_kernel.cu:5-18 -> (n,c); accumulates into `output` when given"""
    input, idx, weight = _f32(input), _i32(idx), _f32(weight)
    n, k = idx.shape
    c = input.shape[1]
    out = np.zeros((n, c), dtype=np.float32) if output is None else _f32(output).copy()
    _load().oracle_pop_interpolation_fwd(n, c, k, _p(input), _p(idx), _p(weight), _p(out))
    return out


def pop_interpolation_bwd(grad_output, idx, weight, m):
    """interpolation_cuda_kernel.cu:20-32 -> grad_input (m,c)"""
    grad_output, idx, weight = _f32(grad_output), _i32(idx), _f32(weight)
    n, k = idx.shape
    c = grad_output.shape[1]
    g = np.zeros((int(m), c), dtype=np.float32)
    _load().oracle_pop_interpolation_bwd(n, c, k, _p(grad_output), _p(idx), _p(weight), _p(g))
    return g


def pop_subtraction_fwd(input1, input2, idx):
    """subtraction_cuda_kernel.cu:5-16 -> (n,nsample,c)"""
    input1, input2, idx = _f32(input1), _f32(input2), _i32(idx)
    n, c = input1.shape
    nsample = idx.shape[1]
    out = np.zeros((n, nsample, c), dtype=np.float32)
    _load().oracle_pop_subtraction_fwd(n, nsample, c, _p(input1), _p(input2), _p(idx), _p(out))
    return out


def pop_subtraction_bwd(idx, grad_output, n2=None):
    """subtraction_cuda_kernel.cu:18-31 -> (grad_input1 (n,c), grad_input2 (n,c))"""
    idx, grad_output = _i32(idx), _f32(grad_output)
    n, nsample, c = grad_output.shape
    g1 = np.zeros((n, c), dtype=np.float32)
    g2 = np.zeros((n if n2 is None else int(n2), c), dtype=np.float32)
    _load().oracle_pop_subtraction_bwd(n, nsample, c, _p(idx), _p(grad_output), _p(g1), _p(g2))
    return g1, g2


def pop_aggregation_fwd(input, position, weight, idx):
    """aggregation_cuda_kernel.cu:5-21 -> (n,c)"""
    input, position, weight, idx = _f32(input), _f32(position), _f32(weight), _i32(idx)
    n, nsample, c = position.shape
    w_c = weight.shape[-1]
    out = np.zeros((n, c), dtype=np.float32)
    _load().oracle_pop_aggregation_fwd(n, nsample, c, w_c, _p(input), _p(position), _p(weight), _p(idx), _p(out))
    return out


def pop_aggregation_bwd(input, position, weight, idx, grad_output):
    """aggregation_cuda_kernel.cu:24-41 -> (grad_input, grad_position, grad_weight)"""
    input, position, weight, idx = _f32(input), _f32(position), _f32(weight), _i32(idx)
    grad_output = _f32(grad_output)
    n, nsample, c = position.shape
    w_c = weight.shape[-1]
    gi = np.zeros_like(input)
    gp = np.zeros_like(position)
    gw = np.zeros_like(weight)
    _load().oracle_pop_aggregation_bwd(n, nsample, c, w_c, _p(input), _p(position), _p(weight), _p(idx),
                                       _p(grad_output), _p(gi), _p(gp), _p(gw))
    return gi, gp, gw
