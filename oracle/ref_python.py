"""The REFERENCE's own Python (openpoints/...) made importable for the parity tests.

TEST INFRASTRUCTURE ONLY — nothing under amcontrast3d_b200/ may import this.

The reference is Python without a setup.py, so "installing" it is making its package importable:
`stage()` copies the .py files of /root/reference/openpoints, unmodified, into baseline/_ref/openpoints
(git-ignored, NOT gpurun-ignored: it travels to the GPU box the way a `pip install --target baseline/_ref`
of a packaged reference would).  Nothing of it enters the repository's history.  `root()` prefers the
checkout where it exists (the build container) and falls back to the staged copy (the GPU box).

`import_reference(tier)` puts the path on sys.path, stubs the third-party packages the reference imports
but the hot path never touches (wandb, easydict, ...; SURVEY.md §8c tier O2), runs
`amcontrast3d_b200.compat.install(tier)` so that `import pointops_cuda` / `import pointnet2_batch_cuda`
inside the reference resolve to the sm_100a library, and imports `openpoints.models` first (the
reference has a circular import otherwise).  On a CUDA device the reference's hard-coded
`.cuda()` / `torch.cuda.IntTensor` calls work as they stand, so — unlike tests/golden/make_loss_golden.py,
which runs it on CPU — nothing of torch is patched here.
"""
from __future__ import annotations

import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CHECKOUT = "/root/reference"
STAGED = os.path.join(REPO, "baseline", "_ref")


def stage(force: bool = False) -> str | None:
    """Copy the reference's Python package into baseline/_ref (only where the checkout exists)."""
    src = os.path.join(CHECKOUT, "openpoints")
    if not os.path.isdir(src):
        return STAGED if os.path.isdir(os.path.join(STAGED, "openpoints")) else None
    dst = os.path.join(STAGED, "openpoints")
    if os.path.isdir(dst) and not force:
        return STAGED
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=lambda d, names: [n for n in names
                                                         if not (n.endswith(".py") or os.path.isdir(os.path.join(d, n)))])
    return STAGED


def root() -> str | None:
    if os.path.isdir(os.path.join(CHECKOUT, "openpoints")):
        return CHECKOUT
    if os.path.isdir(os.path.join(STAGED, "openpoints")):
        return STAGED
    return None


def available() -> bool:
    return root() is not None


_STUBS = ["wandb", "shortuuid", "termcolor", "h5py", "easydict", "multimethod", "torch_scatter", "pyvista", "deepspeed",
          "torcheval", "torcheval.metrics"]


def _stub_third_party():
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
                continue
            except Exception:
                pass
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    if not hasattr(sys.modules["termcolor"], "colored"):
        sys.modules["termcolor"].colored = lambda s, *a, **k: s
    if not hasattr(sys.modules["easydict"], "EasyDict"):
        sys.modules["easydict"].EasyDict = dict
    if not hasattr(sys.modules["multimethod"], "multimethod"):
        sys.modules["multimethod"].multimethod = lambda f: f
    if not hasattr(sys.modules["torcheval.metrics"], "R2Score"):
        sys.modules["torcheval.metrics"].R2Score = object


_imported = None


def import_reference(tier: int = 1):
    """-> namespace of the reference's hot-path callables, running over compat.install(tier)."""
    global _imported
    if _imported is not None:
        return _imported
    r = root()
    if r is None:
        raise RuntimeError("the reference's Python is neither at /root/reference nor staged in baseline/_ref "
                           "(run oracle.ref_python.stage() in the build container)")
    if r not in sys.path:
        sys.path.insert(0, r)
    _stub_third_party()
    from amcontrast3d_b200 import compat
    compat.install(tier=tier)
    import openpoints.models  # noqa: F401  (first: circular import otherwise)
    from openpoints.AMContrast3D import MarginContrast, MaskedRefine, metrics
    from openpoints.AMContrast3D.AEF import ambiguity, utils
    from openpoints.cpp.pointops.functions import pointops
    from openpoints.models.layers import group, subsample, upsampling
    _imported = types.SimpleNamespace(root=r, MarginContrast=MarginContrast, MaskedRefine=MaskedRefine, metrics=metrics,
                                      ambiguity=ambiguity, utils=utils, pointops=pointops, group=group,
                                      subsample=subsample, upsampling=upsampling)
    return _imported


class Args(dict):
    """ambiguity_args as a plain attr-dict (EasyConfig's overloads collapse under the multimethod stub)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__
