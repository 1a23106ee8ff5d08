"""Build libamc3d (the sm_100a CUDA kernels + C-ABI, include/amc3d.h) in-tree with nvcc.

The library links only against the CUDA runtime (static) — no torch headers, no pybind — so
it compiles in seconds and the same .so is what a cgo/JNI/ctypes binding would load
(INTEGRATION.md).  The built file lives in amcontrast3d_b200/lib/ so that it travels with
the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBNAME = "libamc3d_sm100a.so"
LIBPATH = os.path.join(LIBDIR, LIBNAME)

SOURCES = ["common.cu", "knn.cu", "knn_grid.cu", "batch_query.cu", "fps.cu", "group.cu", "amloss.cu", "refine.cu",
           "pointops_packed.cu", "fused_sa.cu", "voxel.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libamc3d (sm_100a)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()


def source_digest() -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "amc3d.h"))
    return _digest(deps)


def is_current() -> bool:
    stamp = LIBPATH + ".sha256"
    return os.path.exists(LIBPATH) and os.path.exists(stamp) and open(stamp).read().strip() == source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libamc3d_sm100a.so.  Returns its path."""
    if not force and is_current():
        return LIBPATH
    nvcc = _nvcc()
    os.makedirs(OBJDIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIBPATH, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(LIBPATH + ".sha256", "w") as fh:
        fh.write(source_digest())
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
