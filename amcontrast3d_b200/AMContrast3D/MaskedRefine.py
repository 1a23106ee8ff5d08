"""Mirror of openpoints/AMContrast3D/MaskedRefine.py: RefinementMethod (:7-119) with the same
constructor and methods.  DualMasks (fusion 'MIN', the shipped setting) runs on the sm_100a
kernels; the remaining small variants are the reference's torch expressions."""
import torch
from torch.functional import F

from .. import _amloss


class LazyRate:
    """Percentage of refined points, computed from a device counter only when somebody looks at it
    (float(), comparison, arithmetic, formatting) — so DualMasks itself never synchronises with the host
    and can be captured in a CUDA graph."""

    def __init__(self, count, numel):
        self.count, self.numel = count, numel

    def __float__(self):
        return (int(self.count.item()) / self.numel) * 100

    def __eq__(self, other):
        return float(self) == float(other)

    def __lt__(self, other):
        return float(self) < float(other)

    def __add__(self, other):
        return float(self) + float(other)

    __radd__ = __add__

    def __mul__(self, other):
        return float(self) * float(other)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return float(self) / float(other)

    def __format__(self, spec):
        return format(float(self), spec)

    def __repr__(self):
        return repr(float(self))


class RefinementMethod():

    def __init__(self, stage_list, p, f, a, i, B, K, fusion, threshold_max, threshold, gamma):
        self.stage_list = stage_list
        self.position = p          # (B,n,3)
        self.feature = f           # (B,D,n)
        self.ambiguity = a         # (B,1,n)
        self.i = i
        self.batch = B
        self.sample_k = K
        self.fusion = fusion
        self.threshold_max = threshold_max
        self.threshold = threshold
        self.gamma = gamma

    def _map(self):
        dim = self.stage_list['ambiguity_map'][self.i].shape[1]
        return self.stage_list['ambiguity_map'][self.i].unsqueeze(0).view(self.batch, dim, -1)

    def MapAttention(self):
        raise NotImplementedError("MapAttention needs the APM Attention block (linear_mapping: True), which "
                                  "no shipped config enables and which is outside the B200 hot path")

    def MapSum(self):
        self.feature = self.feature + self._map()
        return self.feature

    def MapMultiply(self):
        self.feature = torch.mul(self.feature, self._map())
        return self.feature

    def Multiply(self):
        self.feature = torch.mul(self.feature, self.ambiguity)
        return self.feature

    def DualMasks(self):
        """kNN(K) over the flattened batch, per point the neighbour of minimum ambiguity, masked
        chunk replacement, gamma blend (MaskedRefine.py:49-86; bug-compatible chunk view,
        SURVEY.md App. A.6) -> (feature (B,D,n), update rate in %)."""
        xyz = self.position.reshape(-1, 3).contiguous().float()
        o = torch.full((1,), xyz.shape[0], dtype=torch.int32, device=xyz.device)   # no host->device copy
        knn_idx, _ = _amloss.knn_raw(self.sample_k, xyz, xyz, o, o)
        self.sample_k -= 1                                     # the reference mutates it too (:59)
        f = self.feature.contiguous()
        if f.dtype != torch.float32:
            f = f.float()
        a = self.ambiguity.contiguous().float()
        if self.fusion == 'MIN':
            nl = _amloss.NeighbourList(knn_idx, drop_self=True)
            jmin = _amloss.refine_select(nl, a.view(-1))
            count = torch.zeros((1,), dtype=torch.int32, device=f.device)
            self.feature = _amloss.DualMasksFunction.apply(f, a, jmin, self.threshold, self.threshold_max,
                                                           self.gamma, count)
            # the reference returns a Python float here (a host synchronisation per decoder stage); the
            # value is only logged, so it is handed back lazily: float(rate) / formatting reads it
            return self.feature, LazyRate(count, a.numel())
        elif self.fusion == 'MIN_ALL0':
            D = f.shape[1]
            nidx = knn_idx[:, 1:].reshape(-1).long()
            m = knn_idx.shape[0]
            nf = f.view(-1, D)[nidx].view(m, self.sample_k, D)
            na = a.view(-1, 1)[nidx].view(m, self.sample_k, 1)
            cross = torch.mean(nf * ~na.gt(0), dim=1).view(f.shape[0], D, -1)
            mask, rate = self.self_mask()
            f_new = f * ~mask + cross * mask
            self.feature = self.gamma * f_new + (1 - self.gamma) * f
            return self.feature, rate
        raise ValueError(f'unknown fusion {self.fusion!r}')

    def self_mask(self):
        """threshold <= a <= threshold_max (MaskedRefine.py:112-119) -> (mask (B,1,n), rate %)"""
        mask = self.ambiguity.le(self.threshold_max) * self.ambiguity.ge(self.threshold)
        rate = (torch.count_nonzero(mask.long()).item() / self.ambiguity.numel()) * 100
        return mask, rate

    def consistency_regularization(self, embedding_1, embedding_2):
        """Jensen-Shannon divergence (MaskedRefine.py:122-131)"""
        p1 = F.softmax(embedding_1, dim=0)
        p2 = F.softmax(embedding_2, dim=0)
        mid = 0.5 * (p1 + p2)
        loss = F.kl_div(F.log_softmax(embedding_1, dim=0), mid, reduction="batchmean")
        loss += F.kl_div(F.log_softmax(embedding_2, dim=0), mid, reduction="batchmean")
        return 0.5 * loss
