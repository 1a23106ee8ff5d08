"""Mirror of the functions of openpoints/AMContrast3D/AEF/utils.py that the loss uses:
get_subscene_label_CBL (:11-43), fetch_pxo (:46-52), get_ftype (:107-116)."""
import torch
import torch.nn.functional as F

from ... import _amloss


def get_ftype(ftype):
    """utils.py:107-116"""
    if ftype in ['out', 'fout', 'f_out', 'latent', 'logits', 'probs']:
        ptype = 'p_out'
        ftype = 'f_out' if ftype in ['out', 'fout'] else ftype
    elif ftype in ['sample', 'fsample', 'f_sample']:
        ptype = 'p_sample'
        ftype = 'f_sample' if ftype in ['sample', 'fsample'] else ftype
    else:
        raise KeyError(f'not supported ftype = {ftype}')
    return ftype, ptype


def fetch_pxo(stage_n, stage_i, stage_list, ftype):
    """utils.py:46-52 -> (p_out (M,3), f_out (M,D), offset i32)"""
    stage = stage_list[stage_n][stage_i]
    return stage['p_out'], stage['f_out'], stage['offset']


def stage_label_ids(stage_n, stage_i, stage_list, target, nstride, num_classes, ignore_index):
    """Integer form of get_subscene_label_CBL followed by the argmax the loss takes of it
    (MarginContrast.py:112-113): cls (m) i32, and ncls.  This is what the fused path uses."""
    if stage_i == 0:
        return _amloss.stage_labels(target, num_classes, ignore_index, None)
    kr = int(torch.prod(nstride[:stage_i]))
    stage_from = stage_list['up'][0]
    stage_to = stage_list[stage_n][stage_i]
    nidx, _ = _amloss.knn_raw(kr, stage_from['p_out'], stage_to['p_out'], stage_from['offset'], stage_to['offset'])
    return _amloss.stage_labels(target, num_classes, ignore_index, nidx)


def get_subscene_label_CBL(stage_n, stage_i, stage_list, target, nstride, num_classes, ignore_index):
    """utils.py:11-43 -> (m, ncls) float32: one-hot of the target at stage 0, mean one-hot over
    the kr = prod(nstride[:i]) nearest stage-0 points otherwise (ScanNet's ignore_index becomes
    the extra class `num_classes`).  Kept with the reference's return type for callers that
    want the soft label; the kNN runs on the sm_100a kernel."""
    if ignore_index is not None:
        num_classes = num_classes + 1
        target = torch.where(target == ignore_index, torch.full_like(target, num_classes - 1), target)
    x = F.one_hot(target, num_classes)
    if stage_i == 0:
        return x.float()
    kr = int(torch.prod(nstride[:stage_i]))
    stage_from = stage_list['up'][0]
    stage_to = stage_list[stage_n][stage_i]
    nidx, _ = _amloss.knn_raw(kr, stage_from['p_out'], stage_to['p_out'], stage_from['offset'], stage_to['offset'])
    x = x[nidx.view(-1).long(), :].view(stage_to['p_out'].shape[0], kr, x.shape[1])
    return x.float().mean(-2)
