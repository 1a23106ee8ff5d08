"""Mirror of openpoints/AMContrast3D/metrics.py: the evaluation-time ambiguity reporting over a whole room.

posmask_searching (:160-184) is the GPU-heavy part — one kNN over the room as a single segment plus the
label compare; ambiguity_metrics (:33-157) turns the per-point ambiguity into per-bucket accuracy,
confusion matrices and per-class shares; ambiguity_summary (:9-29) averages those over the rooms.
Same signatures, return values and printed lines as the reference.  vis_tsne (plotting) is not mirrored.
"""
import numpy as np
import torch

from .. import _amloss
from .AEF.ambiguity import ambiguity_function


def posmask_searching(xyz, target, nsample, num_classes, ignore_index):
    """xyz (n,3), target (n) i64 -> (posmask (n,nsample-1) bool, neighbor_idx (n,nsample-1) i32)"""
    xyz = xyz.contiguous().float()
    o = torch.full((1,), xyz.shape[0], dtype=torch.int32, device=xyz.device)
    cls, _ = _amloss.stage_labels(target, num_classes, ignore_index, None)
    knn_idx, _ = _amloss.knn_raw(nsample, xyz, xyz, o, o)
    nl = _amloss.NeighbourList(knn_idx, drop_self=True)
    posbits, _, _ = _amloss.posmask_count(nl, cls)
    return _amloss.unpack_posmask(posbits, nl.ke), knn_idx[..., 1:].contiguous()


def _mious(tp, union, count):
    """openpoints/utils/metrics.py:176-183 get_mious, first three values (FP32 like the reference)"""
    tp, union, count = tp.cpu(), union.cpu(), count.cpu()
    iou = (tp + 1e-10) / (union + 1e-10) * 100
    acc = (tp + 1e-10) / (count + 1e-10) * 100
    return torch.mean(iou).item(), torch.mean(acc).item(), (tp.sum() / count.sum() * 100).item()


def _share(part, whole):
    return round(part / whole * 100, 2)


def ambiguity_metrics(p, label, pred, posmask_test, nsample_test, neighbor_idx_test, cctype, ccbeta, vis,
                      cm_0, cm_low, cm_semi, cm_high, cm_1, nu):
    """-> (ambiguity_soft (n), ratio {bucket: accuracy}, ambiguity_count [5 %], [1.0]*5,
    cls {class: [5 %]}, miou[5], macc[5], oa[5], count[5][num_classes])  — metrics.py:33-157.

    The ambiguity a in [0,1] of every point is binned as floor(10 a + 1) in 1..11; the five groups are
    a == 0, 0 < a < nu, a == nu, nu < a < 1, a == 1.  cm_* are the caller's five ConfusionMatrix objects
    (they accumulate over rooms).  Everything past the ambiguity itself is derived from two histograms and
    read back in one transfer each, instead of one host synchronisation per bucket and class."""
    a, a_count = ambiguity_function(p, posmask_test, nsample_test, neighbor_idx_test, cctype, ccbeta, vis, nu)
    mapping = torch.floor(a * 10 + 1)
    nu_m = nu * 10 + 1
    groups = (mapping == 1, torch.logical_and(1 < mapping, mapping < nu_m), mapping == nu_m,
              torch.logical_and(nu_m < mapping, mapping < 11), mapping == 11)

    miou, macc, oa, counts = [], [], [], []
    for cm, sel in zip((cm_0, cm_low, cm_semi, cm_high, cm_1), groups):
        cm.update(pred[sel], label[sel])
        mi, ma, o = _mious(cm.tp, cm.union, cm.count)
        miou.append(round(mi, 2))
        macc.append(round(ma, 2))
        oa.append(round(o, 2))
        counts.append(cm.count.tolist())
    print('miou per ambiguity:', miou)
    print('macc per ambiguity:', macc)
    print('oa per ambiguity:', oa)
    print('count per ambiguity:', counts)

    # accuracy per bin: histogram over (bin, prediction correct)
    bins = mapping.long().clamp_(0, 11)
    hit = (pred == label).long()
    acc_hist = torch.bincount(bins * 2 + hit, minlength=24).view(12, 2).tolist()
    ratio = {}
    for b, (wrong, right) in enumerate(acc_hist):
        if wrong + right:
            ratio[float(b)] = right / (wrong + right)

    # share of the five groups inside every class present in the room; like the reference, the class table
    # keeps the S3DIS boundary (bin 6) whatever nu is (metrics.py:141-145)
    classes, inv = torch.unique(label, return_inverse=True)
    table = torch.bincount(inv * 12 + bins, minlength=classes.numel() * 12).view(-1, 12).tolist()
    cls = {}
    for c, row in zip(classes.tolist(), table):
        whole = sum(row)
        cls[c] = [_share(row[1], whole), _share(sum(row[2:6]), whole), _share(row[6], whole),
                  _share(sum(row[7:11]), whole), _share(row[11], whole)]
    print('(%) count per cls: 0, low=(0,0.5), semi=0.5, high=(0.5,1), 1:', cls)
    return a, ratio, a_count, [1.0] * 5, cls, miou, macc, oa, counts


def ambiguity_summary(num_classes, ambiguity_vs_accuracy_list, ambiguity_vs_count_list,
                      ambiguity_vs_accuracy_lowsemihigh_list, ambiguity_vs_cls_list, ambiguity_cm_miou,
                      ambiguity_cm_macc, ambiguity_cm_oa, ambiguity_cm_count):
    """Averages of the per-room reports over the test set, printed (metrics.py:9-29)."""
    def avg(rows, decimals):
        return np.around(np.mean(rows, axis=0), decimals=decimals)

    for c in range(num_classes):
        rows = [room[c] for room in ambiguity_vs_cls_list if c in room]
        print('count per cls: ', c, avg(rows, 3))
    print('count per a_i: 0, low=(0,0.5), semi=0.5, high=(0.5,1), 1:', avg(ambiguity_vs_count_list, 3))
    print('acc per a_i: 0, low=(0,0.5), semi=0.5, high=(0.5,1), 1:', avg(ambiguity_vs_accuracy_lowsemihigh_list, 3))
    print('miou per ambiguity:', avg(ambiguity_cm_miou, 2))
    print('macc per ambiguity:', avg(ambiguity_cm_macc, 2))
    print('oa per ambiguity:', avg(ambiguity_cm_oa, 2))
    per_group = avg(ambiguity_cm_count, 0)
    for name, row in zip(('count-0:   ', 'count-low: ', 'count-semi:', 'count-high:', 'count-1:   '), per_group):
        print(name, row)
