"""Mirror of openpoints/AMContrast3D/MarginContrast.py: AmbiguityHead (:15-52) and ContrastHead
(:56-273) with the reference's constructor, attributes, forward signatures and return values.

Per stage the reference runs ~70 ATen kernels, materialises [m,k-1,ncls] and [m,k-1,D] neighbour
tensors several times over and loops in Python over every boundary point; here a stage is
  kNN (label vote) -> stage_labels -> kNN (self) -> posmask_count -> ambiguity -> fused loss
on the sm_100a kernels (amcontrast3d_b200/csrc/{knn,amloss}.cu), with no host synchronisation.
Both modules stay parameter- and buffer-free so the published checkpoints keep loading
(checkpoint row).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _amloss
from .AEF.utils import fetch_pxo, get_ftype, get_subscene_label_CBL, stage_label_ids
from .AEF.function import _eps
from .AEF.ambiguity import ambiguity_function


def _stage_ambiguity(n, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args, nstride, ftype):
    """Steps 1-4 of SURVEY.md App. A.4 for one stage -> dict(p, features, nl, posbits, cnt, a, stats, knn_idx)."""
    stage = stageACE_list[n][i]
    p, features, o = stage['p_out'], stage.get('f_out'), stage['offset']     # fetch_pxo; features may be absent
    p = p.contiguous()
    if p.dtype != torch.float32:
        p = p.float()
    nsample = int(ambiguity_args.nsample)
    cls, _ = stage_label_ids(n, i, stageACE_list, target, nstride, num_classes, ignore_index)
    knn_idx, _, order = _amloss.knn_raw(nsample, p, p, o, o, want_order=True)
    nl = _amloss.NeighbourList(knn_idx, drop_self=True)       # the reference's [..., 1:] without the copy
    posbits, cnt, max_cnt = _amloss.posmask_count(nl, cls)
    a, stats = _amloss.ambiguity(p, nl, posbits, cnt, max_cnt, ambiguity_args.cctype, ambiguity_args.ccbeta,
                                 ambiguity_args.nu)
    return dict(p=p, features=features, nl=nl, posbits=posbits, cnt=cnt, a=a, stats=stats, knn_idx=knn_idx, cls=cls,
                order=order)


class AmbiguityHead(nn.Module):
    """Returns the per-stage ambiguity a_s only (MarginContrast.py:15-52)."""

    def __init__(self):
        super().__init__()
        self.nstride = torch.tensor([4, 4, 4, 4])
        self.ftype = get_ftype('latent')[0]
        self.posmask_func = self.posmask_cnt
        self.main = self.point_ambiguity

    def posmask_cnt(self, labels, neighbor_label):
        """(m,ncls), (m,k,ncls) soft labels -> (m,k) bool (MarginContrast.py:23-27)"""
        return torch.argmax(torch.unsqueeze(labels, -2), -1) == torch.argmax(neighbor_label, -1)

    def point_ambiguity(self, n, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args):
        return _stage_ambiguity(n, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args,
                                self.nstride, self.ftype)['a']

    def forward(self, target, stageACE_list, num_classes, ignore_index, ambiguity_args):
        return [self.main(ambiguity_args.stages, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args)
                for i in range(ambiguity_args.stages_num)]


class ContrastHead(nn.Module):
    """Adaptive-margin contrastive loss over the decoder stages (MarginContrast.py:56-273).
    forward(...) -> (loss_sum, cat(a_s), [a_s])."""

    def __init__(self):
        super().__init__()
        self.nstride = torch.tensor([4, 4, 4, 4])
        self.stages = [('up', 0), ('up', 1), ('up', 2), ('up', 3)]
        self.ftype = get_ftype('latent')[0]
        self.project = None
        self.dist_func = self.dist_cos
        self.contrast_func = self.contrast_softnn_margin
        self.posmask_func = self.posmask_cnt
        self.main_contrast = self.point_contrast_margin

    # ---- the reference's small building blocks, kept callable (used by the torch-composed path) ----
    def dist_dot(self, features, neighbor_feature):
        return torch.sum(torch.mul(torch.unsqueeze(features, -2), neighbor_feature), -1) + _eps

    def dist_cos(self, features, neighbor_feature):
        return F.cosine_similarity(torch.unsqueeze(features, -2), neighbor_feature, dim=2)

    def dist_l2(self, features, neighbor_feature):
        d = torch.unsqueeze(features, -2) - neighbor_feature
        return torch.sqrt(torch.sum(d ** 2, axis=-1) + _eps)

    def posmask_cnt(self, labels, neighbor_label):
        return torch.argmax(torch.unsqueeze(labels, -2), -1) == torch.argmax(neighbor_label, -1)

    def contrast_softnn_margin(self, dist, posmask, ambiguity, ambiguity_args, invalid_mask=None):
        """torch composition of MarginContrast.py:117-174 for (m,k) similarities; used for the
        option combinations the fused kernel does not cover (margin == 'learned', invalid_mask)."""
        if ambiguity_args.margin == 'constant':
            margin = ambiguity_args.nu
        elif ambiguity_args.margin == 'adaptive':
            margin = ambiguity_args.mu * torch.unsqueeze(ambiguity, -1) + ambiguity_args.nu
        elif ambiguity_args.margin == 'learned':
            u = torch.mean(dist * ~posmask, 1)
            v = torch.mean(dist * posmask, 1)
            margin = (torch.unsqueeze(u, -1) - 1) * torch.unsqueeze(ambiguity, -1) + torch.unsqueeze(v, -1)
        else:
            raise ValueError(f'unknown margin {ambiguity_args.margin!r}')
        if ambiguity_args.db == '-m':
            dist = (dist - margin) * posmask + dist * ~posmask
        elif ambiguity_args.db == '+m':
            dist = dist * posmask + (dist + margin) * ~posmask
        if ambiguity_args.temperature is not None:
            dist = dist / ambiguity_args.temperature
        exp = torch.exp(dist)
        if invalid_mask is not None:
            exp = exp * (1 - invalid_mask)
        pos = torch.sum(exp * posmask, axis=-1)
        neg = torch.sum(exp * (1 - posmask.int()), axis=-1)
        if ambiguity_args.supervisedCL == 'Method1':
            loss = pos / torch.sum(exp, axis=-1) + _eps
        elif ambiguity_args.supervisedCL == 'Method2':
            loss = exp * posmask / (exp * posmask + neg.unsqueeze(-1)) + _eps
            loss = torch.sum(loss, axis=-1) / (torch.sum(posmask.int(), axis=-1) + _eps)
        else:
            raise ValueError(f'unknown supervisedCL {ambiguity_args.supervisedCL!r}')
        return -torch.log(loss)

    def precompute_geometry(self, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args):
        """Everything of stage i that depends on coordinates and labels only — stage labels, kNN, posmask,
        ambiguity (steps 1-4 of SURVEY.md App. A.4) — so that a trainer can run it ahead of time, e.g. on a
        side stream next to the encoder.  `stageACE_list[stages][s]` needs 'p_out' and 'offset' for s = 0
        and s = i only.  Put the returned objects in stageACE_list['am_geometry'] (a list indexed by stage,
        None = compute as usual) and forward() uses them; results are identical."""
        st = _stage_ambiguity(ambiguity_args.stages, i, stageACE_list, target, num_classes, ignore_index,
                              ambiguity_args, self.nstride, self.ftype)
        # off the critical path there is time to drop the anchors the loss will not select from the visiting
        # order, which keeps the forward kernel's warps full (narrow stages put several anchors in a warp)
        st['order'] = _amloss.compact_order(st['order'], st['a'])
        return st

    def point_contrast_margin(self, n, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args):
        """One stage (MarginContrast.py:220-259) -> (loss, output_ai, target_ai)."""
        pre = stageACE_list.get('am_geometry') if hasattr(stageACE_list, 'get') else None
        if pre is not None and pre[i] is not None:
            st = dict(pre[i])
            st['features'] = stageACE_list[n][i]['f_out']
        else:
            st = _stage_ambiguity(n, i, stageACE_list, target, num_classes, ignore_index, ambiguity_args,
                                  self.nstride, self.ftype)
        a = st['a']
        target_ai = torch.clone(a)
        output_ai = stageACE_list['ambiguity'][i].flatten() if 'ambiguity' in stageACE_list.keys() else None
        features = st['features']
        if _amloss.fused_supported(ambiguity_args):
            loss = _amloss.am_loss(features, st['nl'], st['posbits'], a, st['stats'], ambiguity_args, st['order'])
        else:
            # torch composition on device over the same kNN / posmask / ambiguity
            sel = torch.logical_and(0 < a, a <= 1)
            nidx = st['knn_idx'][:, 1:][sel].long()
            posmask = _amloss.unpack_posmask(st['posbits'], st['nl'].ke)[sel]
            dist = self.dist_func(features[sel], features[nidx.reshape(-1)].view(nidx.shape[0], nidx.shape[1], -1))
            loss = torch.mean(self.contrast_func(dist, posmask, a[sel], ambiguity_args))
        return loss, output_ai, target_ai

    def forward(self, output, target, stageACE_list, num_classes, ignore_index, ambiguity_args):
        loss_sum = 0
        target_ai_list = []
        for i in range(ambiguity_args.stages_num):
            loss, output_ai, target_ai = self.main_contrast(ambiguity_args.stages, i, stageACE_list, target,
                                                            num_classes, ignore_index, ambiguity_args)
            loss_sum += loss
            target_ai_list.append(target_ai)
        return loss_sum, torch.cat(target_ai_list), target_ai_list
