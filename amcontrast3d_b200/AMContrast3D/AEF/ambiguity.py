"""Mirror of openpoints/AMContrast3D/AEF/ambiguity.py:11-93."""
import torch

from ... import _amloss


def percentages(stats, m):
    """the reference's [count_0, count_low, count_semi, count_high, count_1] (ambiguity.py:85-91);
    one device->host copy instead of five .item() syncs"""
    s = stats[2:7].tolist()
    return [round(v / m * 100, 2) for v in s]


def ambiguity_function(p, posmask, nsample, neighbor_idx, ambiguity_type, ambiguity_beta, ambiguity_vis, nu):
    """p (m,3), posmask (m,k-1) bool, neighbor_idx (m,k-1) i32 (self column already dropped) ->
    (a (m) f32, [5 percentages]).  Same signature as the reference; `ambiguity_vis` (pyvista
    pop-ups) is not supported on a headless training box and must be False."""
    if ambiguity_vis:
        raise NotImplementedError("ambiguity_vis opens pyvista windows in the reference; not part of the hot path")
    nl = _amloss.NeighbourList(neighbor_idx.contiguous().int(), drop_self=False)
    posbits, cnt = _amloss.pack_posmask(posmask)
    max_cnt = cnt.max().reshape(1).int()
    a, stats = _amloss.ambiguity(p.contiguous().float(), nl, posbits, cnt, max_cnt, ambiguity_type,
                                 ambiguity_beta, nu)
    return a, percentages(stats, a.shape[0])
