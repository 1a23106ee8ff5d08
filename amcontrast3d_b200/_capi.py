"""ctypes binding of libamc3d (include/amc3d.h) — the only door from Python to the kernels.

There is deliberately no CPU fallback: if the shared library is missing or a call fails,
this raises.  Device pointers come from ``tensor.data_ptr()``; the stream is torch's current
stream on the tensor's device, so the kernels order correctly with surrounding torch work
(the reference launches on the legacy default stream).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_longlong, c_void_p

from . import _build

_lock = threading.Lock()
_lib = None


class LossParams(Structure):
    """struct amc3d_loss_params (include/amc3d.h)"""

    _fields_ = [
        ("temperature", c_float),
        ("has_temperature", c_int),
        ("margin_mode", c_int),
        ("mu", c_float),
        ("nu", c_float),
        ("db_mode", c_int),
        ("cl_method", c_int),
    ]


_P = c_void_p
_I = c_int
_F = c_float
_LL = c_longlong

# name -> argtypes (restype is always int unless listed in _RESTYPES)
SIGNATURES = {
    "amc3d_version": [],
    "amc3d_arch": [],
    "amc3d_last_error": [],
    "amc3d_trim_scratch": [ctypes.c_size_t],
    "amc3d_fp32_probe": [_I, _I, _P, POINTER(ctypes.c_double), _P],
    "amc3d_search_stats": [_P],
    "amc3d_furthest_point_sampling": [_I, _I, _I, _P, _P, _P, _P],
    "amc3d_ball_query": [_I, _I, _I, _F, _I, _P, _P, _P, _P],
    "amc3d_group_points": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "amc3d_group_points_ws": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_group_points_grad": [_I, _I, _I, _I, _I, _P, _P, _P, _P],
    "amc3d_group_points_grad_ws": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_group_points_grad_ws_set": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_group_xyz_relative": [_I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P],
    "amc3d_gather_points": [_I, _I, _I, _I, _P, _P, _P, _P],
    "amc3d_gather_points_grad": [_I, _I, _I, _I, _P, _P, _P, _P],
    "amc3d_transpose_batched": [_I, _I, _I, _P, _P, _P],
    "amc3d_fused_sa_forward": [_I, _I, _I, _I, _I, _I, _F, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                               _P, _P, _P, _P],
    "amc3d_fused_sa_backward_scatter": [_I, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P],
    "amc3d_fused_sa_backward_coefs": [_I, _I, ctypes.c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "amc3d_fused_sa_backward_assemble": [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "amc3d_debug_fastdiv": [ctypes.c_uint, ctypes.c_uint],
    "amc3d_fused_sa_moments": [_I, _I, _I, _I, _F, _I, _P, _P, _P, _P, _I, _P, _I, _P, _P],
    "amc3d_voxel_keys": [_LL, ctypes.c_double, _P, _P, _P, _P],
    "amc3d_crop_dist2": [_LL, _P, _LL, _P, _P],
    "amc3d_three_nn": [_I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_three_interpolate": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_three_interpolate_grad": [_I, _I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_three_interpolate_ws": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_three_interpolate_grad_ws": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_three_interpolate_grad_ws_set": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_knnquery": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
    "amc3d_knnquery_order": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "amc3d_grouping_forward": [_I, _I, _I, _P, _P, _P, _P],
    "amc3d_grouping_backward": [_I, _I, _I, _P, _P, _P, _P],
    "amc3d_pointops_ballquery": [_I, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_pointops_furthestsampling": [_I, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_pointops_interpolation_forward": [_I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_pointops_interpolation_backward": [_I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_pointops_subtraction_forward": [_I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_pointops_subtraction_backward": [_I, _I, _I, _P, _P, _P, _P, _P],
    "amc3d_pointops_aggregation_forward": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_pointops_aggregation_backward": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "amc3d_class_counts": [_I, _I, _P, _P, _P, _I, _P, _P],
    "amc3d_stage_labels": [_I, _I, _I, _I, _LL, _P, _P, _P, _P],
    "amc3d_posmask_count": [_I, _I, _I, _P, _P, _P, _P, _P, _P],
    "amc3d_ambiguity": [_I, _I, _I, _P, _P, _P, _P, _P, _I, _F, _F, _P, _P, _P],
    "amc3d_ambiguity_backend": [_I, _I, _I, _P, _P, _P, _P, _P, _I, _F, _F, _I, _P, _P, _P],
    "amc3d_row_inv_norm": [_I, _I, _P, _P, _P],
    "amc3d_amloss_forward": [_I, _I, _I, _I, _P, _P, _P, _P, _P, POINTER(LossParams), _P, _P, _P],
    "amc3d_amloss_forward_order": [_I, _I, _I, _I, _P, _P, _P, _P, _P, POINTER(LossParams), _P, _P, _P, _P],
    "amc3d_amloss_reduce": [_I, _P, _P, _P, _P],
    "amc3d_amloss_backward": [_I, _I, _P, _P, _P, _P, _P, _I, _P, _P],
    "amc3d_refine_select": [_I, _I, _I, _P, _P, _P, _P],
    "amc3d_refine_forward": [_I, _I, _I, _P, _P, _P, _F, _F, _F, _P, _P, _P],
    "amc3d_refine_backward": [_I, _I, _I, _P, _P, _P, _F, _F, _F, _P, _P],
}
_RESTYPES = {"amc3d_arch": c_char_p, "amc3d_last_error": c_char_p, "amc3d_debug_fastdiv": ctypes.c_uint}


class Amc3dError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIBPATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load libamc3d_sm100a.so (building it with nvcc if absent).  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("AMC3D_LIB")                        # AMC3D_LIB: an experimental build of the same ABI
        if not path:
            path = _build.LIBPATH
            # never load a library that was built from other sources than the ones next to it: the digest
            # stamp is written beside the artefact by the build (both untracked), so a checkout that
            # changes csrc/ makes is_current() false until the library is rebuilt
            if not _build.is_current():
                if not build_if_missing:
                    what = "is missing" if not os.path.exists(path) else "is stale (csrc/ or include/amc3d.h changed since it was built)"
                    raise Amc3dError(f"{path} {what}: run `python -m amcontrast3d_b200._build`")
                _build.build()
        elif not os.path.exists(path):
            raise Amc3dError(f"AMC3D_LIB={path} does not exist")
        lib = ctypes.CDLL(path)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = the library does not match the header
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        if lib.amc3d_arch() != b"sm_100a":
            raise Amc3dError("libamc3d was not built for sm_100a")
        _lib = lib
    return _lib


def check(rc: int, name: str) -> None:
    if rc != 0:
        msg = load().amc3d_last_error().decode()
        raise Amc3dError(f"{name} failed (code {rc}): {msg}")


# kernels launched by this package since import (bench.py reports the per-step delta as gpu_launches)
LAUNCHES = 0
# when set to a list, every call appends (name, start_event, stop_event) recorded on the launch stream
PROFILE = None
# measurement hook: when set, called as AFTER_CALL(name, args) after every entry point (bench.py's pair counting)
AFTER_CALL = None


def _kernels_in(name: str, args) -> int:
    if name == "amc3d_group_points_ws":
        return 2 if args[-2] else 1            # transpose + gather with a workspace
    if name in ("amc3d_group_points_grad_ws", "amc3d_group_points_grad_ws_set"):
        return 2 if args[-2] else 1            # scatter + transpose-accumulate (plus a memset)
    if name == "amc3d_three_interpolate_ws":
        return 2 if args[-2] else 1            # transpose + interpolate
    if name in ("amc3d_three_interpolate_grad_ws", "amc3d_three_interpolate_grad_ws_set"):
        return 2 if args[-2] else 1            # scatter + transpose-accumulate (plus a memset)
    if name == "amc3d_fused_sa_forward":
        return 4                               # weight tiles + GEMM + statistics + normalise
    if name in ("amc3d_refine_backward", "amc3d_ambiguity", "amc3d_ambiguity_backend"):
        return 2                               # (boundary count + ambiguity)
    return 1


_FN = {}


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on a non-zero code."""
    global LAUNCHES
    fn = _FN.get(name)
    if fn is None:
        fn = _FN[name] = getattr(load(), name)
    if PROFILE is None:
        rc = fn(*args)
        if rc != 0:
            check(rc, name)
    else:
        import torch

        st = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.default_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        check(fn(*args), name)
        e1.record(st)
        PROFILE.append((name, e0, e1, args[:6]))
    if AFTER_CALL is not None:
        AFTER_CALL(name, args)
    LAUNCHES += _kernels_in(name, args)


def ptr(t) -> int:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return 0
    if not t.is_cuda:
        raise Amc3dError("amc3d kernels need CUDA tensors: there is no CPU path in this package "
                         "(the CPU restatement lives in oracle/ and is test infrastructure only)")
    return t.data_ptr()


_raw_stream = None


def stream(t) -> int:
    """torch's current CUDA stream on the tensor's device, as the raw cudaStream_t value.  Through torch's C
    accessor (one C call, ~0.2 us) rather than torch.cuda.current_stream(...).cuda_stream (a Python object per
    call, ~4 us): with ~170 launches per step that alone is 0.7 ms of an eager step."""
    global _raw_stream
    if _raw_stream is None:
        import torch

        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (
            lambda idx: torch.cuda.current_stream(idx).cuda_stream)
    return _raw_stream(t.device.index)


_NO_CPU = ("amc3d kernels need CUDA tensors: there is no CPU path in this package "
           "(the CPU restatement lives in oracle/ and is test infrastructure only)")


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()
_cur_dev = None


def guard(t):
    """Device guard for the launch; refuses CPU tensors loudly.  Switching devices is only needed
    when the tensor does not live on the current one (never, with one process per GPU)."""
    if not t.is_cuda:
        raise Amc3dError(_NO_CPU)
    global _cur_dev
    if _cur_dev is None:
        import torch

        _cur_dev = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device
    if t.device.index == _cur_dev():
        return _NO_GUARD
    import torch

    return torch.cuda.device(t.device)
