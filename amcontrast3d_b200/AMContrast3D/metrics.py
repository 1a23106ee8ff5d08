"""Mirror of posmask_searching, openpoints/AMContrast3D/metrics.py:160-184 (the evaluation-time
kNN + label compare over a whole room).  The rest of that file is accuracy reporting."""
import torch

from .. import _amloss


def posmask_searching(xyz, target, nsample, num_classes, ignore_index):
    """xyz (n,3), target (n) i64 -> (posmask (n,nsample-1) bool, neighbor_idx (n,nsample-1) i32)"""
    xyz = xyz.contiguous().float()
    o = torch.full((1,), xyz.shape[0], dtype=torch.int32, device=xyz.device)
    cls, _ = _amloss.stage_labels(target, num_classes, ignore_index, None)
    knn_idx, _ = _amloss.knn_raw(nsample, xyz, xyz, o, o)
    nl = _amloss.NeighbourList(knn_idx, drop_self=True)
    posbits, _, _ = _amloss.posmask_count(nl, cls)
    return _amloss.unpack_posmask(posbits, nl.ke), knn_idx[..., 1:].contiguous()
