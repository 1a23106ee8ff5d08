"""Mirror of the two AMContrast3D criteria of openpoints/loss/build.py: CrossEntropyAce (:324-346)
and CrossEntropyAcePre (:281-319), plus a minimal LOSS registry so cfg `NAME`s resolve."""
import torch
import torch.nn as nn

from .AMContrast3D.MarginContrast import ContrastHead

LOSS = {}


def register(cls):
    LOSS[cls.__name__] = cls
    return cls


@register
class CrossEntropyAce(nn.Module):
    """w1 * CE(logit, target) + w2 * AM-contrast (build.py:324-346)"""

    def __init__(self, **kwargs):
        super().__init__()
        self.creterion = nn.CrossEntropyLoss()
        self.contrast_head = ContrastHead()

    def forward(self, logit, target, stageACE_list, num_classes, ignore_index, ambiguity_args):
        logit = logit.transpose(1, 2).reshape(-1, logit.shape[1])
        target = target.flatten()
        ce = self.creterion(logit, target)
        am, _, _ = self.contrast_head(logit, target, stageACE_list, num_classes, ignore_index, ambiguity_args)
        return ambiguity_args.w1 * ce + ambiguity_args.w2 * am


@register
class CrossEntropyAcePre(nn.Module):
    """(w1*CE + w2*AM, w1*CE, w2*AM, w3*L1(APM a, target a)) (build.py:281-319)"""

    def __init__(self, **kwargs):
        super().__init__()
        self.creterion = nn.CrossEntropyLoss()
        self.contrast_head = ContrastHead()
        self.MAE = nn.L1Loss()
        self.MSE = nn.MSELoss()
        self.HUBER = nn.HuberLoss(reduction='mean', delta=0.1)

    def forward(self, logit, target, stageACE_list, num_classes, ignore_index, ambiguity_args):
        logit = logit.transpose(1, 2).reshape(-1, logit.shape[1])
        target = target.flatten()
        ce = self.creterion(logit, target)
        am, target_ai, _ = self.contrast_head(logit, target, stageACE_list, num_classes, ignore_index,
                                              ambiguity_args)
        logits_ai = torch.cat(stageACE_list['ambiguity']).flatten()
        reg = self.MAE(logits_ai, target_ai)
        ce = ambiguity_args.w1 * ce
        am = ambiguity_args.w2 * am
        reg = ambiguity_args.w3 * reg
        return ce + am, ce, am, reg


def build_criterion_from_cfg(cfg, **kwargs):
    """build.py:348-357: cfg.NAME selects the criterion"""
    cfg = dict(cfg)
    return LOSS[cfg.pop('NAME')](**cfg, **kwargs)
