"""Put this package's implementations on the reference's import paths.

Three levels, matching the tiers of the drop-in boundary:

    install(tier=1)   the two compiled extension modules only.  ``import pointops_cuda`` and
                      ``import pointnet2_batch_cuda`` (what openpoints/cpp/pointnet2_batch/__init__.py:2
                      and openpoints/cpp/pointops/functions/pointops.py:7 do) resolve to the
                      sm_100a library; the reference's own Python wrappers, autograd Functions and
                      loss code keep running on top of it, unchanged.
    install(tier=2)   additionally rebinds the operator callables inside the reference's
                      openpoints.models.layers.{subsample,group,upsampling} and
                      openpoints.cpp.pointops.functions.pointops modules (if importable) to the
                      ones of this package (workspace-aware grouping, no extra zero-fills).
    install(tier=3)   additionally rebinds ContrastHead / AmbiguityHead / ambiguity_function /
                      get_subscene_label_CBL / RefinementMethod / posmask_searching and the two
                      criteria, i.e. the fused loss path.

    install(tier=4)   additionally routes LocalAggregation.forward / SetAbstraction.forward of the reference's
                      PointNeXt backbones (openpoints.models.backbone.pointnext_AA / pointnext_MM) through the
                      fused grouping -> conv -> BatchNorm -> ReLU -> max operator (layers/fused.py); layers the
                      operator does not cover keep the reference's composition.

Call it once, before the reference's model / criterion modules are imported (tier 1) or right after
(tiers 2-3 patch attributes of already-imported modules as well).  Nothing here touches files of the
reference checkout.
"""
from __future__ import annotations

import importlib
import sys

_T2 = {
    "openpoints.models.layers.subsample": ("amcontrast3d_b200.layers.subsample",
                                           ["furthest_point_sample", "gather_operation", "FurthestPointSampling"]),
    "openpoints.models.layers.group": ("amcontrast3d_b200.layers.group",
                                       ["grouping_operation", "ball_query", "GroupingOperation", "BallQuery",
                                        "QueryAndGroup", "create_grouper"]),
    "openpoints.models.layers.upsampling": ("amcontrast3d_b200.layers.upsampling",
                                            ["three_nn", "three_interpolate", "three_interpolation", "ThreeNN",
                                             "ThreeInterpolate"]),
    "openpoints.cpp.pointops.functions.pointops": ("amcontrast3d_b200.pointops",
                                                   ["knnquery", "KNNQuery", "furthestsampling", "FurthestSampling",
                                                    "ballquery", "BallQuery", "grouping", "Grouping", "querygroup",
                                                    "queryandgroup", "subtraction", "Subtraction", "aggregation",
                                                    "Aggregation", "interpolation", "interpolation2",
                                                    "Interpolation"]),
}
_T3 = {
    "openpoints.AMContrast3D.MarginContrast": ("amcontrast3d_b200.AMContrast3D.MarginContrast",
                                               ["ContrastHead", "AmbiguityHead"]),
    "openpoints.AMContrast3D.AEF.ambiguity": ("amcontrast3d_b200.AMContrast3D.AEF.ambiguity", ["ambiguity_function"]),
    "openpoints.AMContrast3D.AEF.utils": ("amcontrast3d_b200.AMContrast3D.AEF.utils", ["get_subscene_label_CBL"]),
    "openpoints.AMContrast3D.MaskedRefine": ("amcontrast3d_b200.AMContrast3D.MaskedRefine", ["RefinementMethod"]),
    "openpoints.AMContrast3D.metrics": ("amcontrast3d_b200.AMContrast3D.metrics", ["posmask_searching"]),
    "openpoints.loss.build": ("amcontrast3d_b200.loss", ["CrossEntropyAce", "CrossEntropyAcePre"]),
}


def _rebind(table, strict):
    done = []
    for ref_name, (our_name, attrs) in table.items():
        try:
            ref_mod = sys.modules.get(ref_name) or importlib.import_module(ref_name)
        except Exception:
            if strict:
                raise
            continue
        ours = importlib.import_module(our_name)
        for a in attrs:
            if hasattr(ours, a):
                setattr(ref_mod, a, getattr(ours, a))
                done.append(f"{ref_name}.{a}")
        # the reference re-exports the layer callables from openpoints.models.layers (layers/__init__.py:10-12)
        parent = sys.modules.get(ref_name.rsplit(".", 1)[0])
        if parent is not None:
            for a in attrs:
                if hasattr(parent, a) and hasattr(ours, a):
                    setattr(parent, a, getattr(ours, a))
    return done


def install(tier: int = 1, strict: bool = False):
    """See the module docstring.  Returns the list of names that were (re)bound."""
    from . import _capi, pointnet2_batch_cuda, pointops_cuda

    _capi.load()                               # fail now, loudly, if the sm_100a library is missing
    sys.modules["pointnet2_batch_cuda"] = pointnet2_batch_cuda
    sys.modules["pointops_cuda"] = pointops_cuda
    done = ["pointnet2_batch_cuda", "pointops_cuda"]
    # modules of the reference that were imported before install() hold a reference to the old extension
    for name, attr, mod in (("openpoints.cpp.pointnet2_batch", "pointnet2_cuda", pointnet2_batch_cuda),
                            ("openpoints.cpp", "pointnet2_cuda", pointnet2_batch_cuda),
                            ("openpoints.models.layers.subsample", "pointnet2_cuda", pointnet2_batch_cuda),
                            ("openpoints.models.layers.group", "pointnet2_cuda", pointnet2_batch_cuda),
                            ("openpoints.models.layers.upsampling", "pointnet2_cuda", pointnet2_batch_cuda),
                            ("openpoints.cpp.pointops.functions.pointops", "pointops_cuda", pointops_cuda)):
        m = sys.modules.get(name)
        if m is not None and hasattr(m, attr):
            setattr(m, attr, mod)
            done.append(f"{name}.{attr}")
    if tier >= 2:
        done += _rebind(_T2, strict)
    if tier >= 3:
        done += _rebind(_T3, strict)
    if tier >= 4:
        from .layers import fused
        for name in ("openpoints.models.backbone.pointnext_AA", "openpoints.models.backbone.pointnext_MM",
                     "openpoints.models.backbone.pointnext"):
            try:
                mod = sys.modules.get(name) or importlib.import_module(name)
            except Exception:
                if strict:
                    raise
                continue
            for cls_name, which in (("LocalAggregation", "la"), ("SetAbstraction", "sa")):
                cls = getattr(mod, cls_name, None)
                if cls is not None:
                    fused.bind(cls, which)
                    done.append(f"{name}.{cls_name}.forward")
    return done
