"""Mirror of the voxel down-sampling and cropping of openpoints/dataset/data_util.py on CUDA tensors
(SURVEY.md §8f rank 4): `voxelize` (:125-143, with `fnv_hash_vec` :92-105 / `ravel_hash_vec` :108-122) and
`crop_pc` (:137-174).  Same names, arguments, defaults and return conventions; inputs and outputs are torch CUDA
tensors instead of numpy arrays, so a batch is prepared where it already lives.

Determinism contract.  The reference is not deterministic by itself: `voxelize` sorts its keys with numpy's default
(unstable) argsort and, in train mode, draws from numpy's GLOBAL Mersenne-Twister stream; `crop_pc` draws its seed
point and its shuffle from the same stream.  This implementation fixes what the reference leaves open and makes the
randomness explicit:
  * voxel membership (keys), the enumeration of voxels (ascending key, as `np.unique`) and the counts are exactly the
    reference's;
  * inside a voxel the points are in ascending original index (a STABLE sort) — the reference's order there is
    whatever its unstable sort produced;
  * train mode selects member `rand[v] % count[v]` of voxel v; `rand` is an argument (the reference's
    `np.random.randint(0, count.max(), count.size)` can be passed in), or comes from `generator` (torch's
    counter-based Philox stream);
  * `crop_pc` takes `init_idx` and `shuffle_perm` the same way.
With the same draws the outputs equal the reference's whenever the reference's argsort happens to be stable on the
input (always for voxels of one point), and always as SETS per voxel.  tests/test_gpu_data_util.py checks exactly
that against fixtures produced by the reference's own functions.
"""
from __future__ import annotations

import torch

from . import _capi
from ._capi import ptr, stream

_SIGN = -(1 << 63)          # flips the sign bit: unsigned order of uint64 keys under torch's signed int64 sort


def _keys(coord: torch.Tensor, voxel_size: float, want_cells: bool = False):
    c = coord.detach().contiguous().float()
    n = c.shape[0]
    keys = torch.empty((n,), dtype=torch.int64, device=c.device)
    cells = torch.empty((n, 3), dtype=torch.int64, device=c.device) if want_cells else None
    with _capi.guard(c):
        _capi.call("amc3d_voxel_keys", n, float(voxel_size), ptr(c), ptr(keys), ptr(cells), stream(c))
    return keys, cells


def fnv_hash_vec(cells: torch.Tensor) -> torch.Tensor:
    """FNV64-1A over the columns of an integer array (data_util.py:92-105); int64 bit patterns of the uint64 hashes"""
    h = torch.full((cells.shape[0],), 14695981039346656037 - (1 << 64), dtype=torch.int64, device=cells.device)
    for j in range(cells.shape[1]):
        h = h * 1099511628211
        h = torch.bitwise_xor(h, cells[:, j].long())
    return h


def ravel_hash_vec(cells: torch.Tensor) -> torch.Tensor:
    """Fortran-style ravel of the cell coordinates after subtracting their minimum (data_util.py:108-122)"""
    a = cells.long() - cells.long().min(0)[0]
    amax = a.max(0)[0] + 1
    keys = torch.zeros((a.shape[0],), dtype=torch.int64, device=a.device)
    for j in range(a.shape[1] - 1):
        keys = (keys + a[:, j]) * amax[j + 1]
    return keys + a[:, -1]


def voxelize(coord, voxel_size=0.05, hash_type='fnv', mode=0, rand=None, generator=None):
    """mode 0 (train): one point index per occupied voxel, voxels in ascending key order.
    mode 1 (val): (idx_sort, voxel_idx, count) as the reference returns them."""
    if hash_type == 'ravel':
        _, cells = _keys(coord, voxel_size, want_cells=True)
        key = ravel_hash_vec(cells)
        order_key = key
    else:
        key, _ = _keys(coord, voxel_size)
        order_key = torch.bitwise_xor(key, torch.tensor(_SIGN, dtype=torch.int64, device=key.device))
    key_sort, idx_sort = torch.sort(order_key, stable=True)
    _, voxel_idx, count = torch.unique_consecutive(key_sort, return_inverse=True, return_counts=True)
    if mode == 0:
        start = torch.cumsum(count, 0) - count
        if rand is None:
            rand = torch.randint(0, int(count.max()), (count.numel(),), device=count.device, generator=generator)
        rand = torch.as_tensor(rand, device=count.device).long()
        return idx_sort[start + rand % count]
    return idx_sort, voxel_idx, count


def crop_pc(coord, feat, label, split='train', voxel_size=0.04, voxel_max=None, downsample=True, variable=True,
            shuffle=True, rand=None, init_idx=None, shuffle_perm=None, padding_choice=None, generator=None):
    """-> (coord f32 shifted to its minimum, feat f32 or None, label i64 or None), data_util.py:137-174."""
    if voxel_size and downsample:
        coord = coord - coord.min(0)[0]
        uniq = voxelize(coord, voxel_size, rand=rand, generator=generator)
        coord = coord[uniq]
        feat = feat[uniq] if feat is not None else None
        label = label[uniq] if label is not None else None
    if voxel_max is not None:
        crop_idx = None
        N = coord.shape[0]
        if N >= voxel_max:
            if init_idx is None:
                init_idx = (int(torch.randint(N, (1,), device=coord.device, generator=generator)) if 'train' in split
                            else N // 2)
            c = coord.detach().contiguous().float()
            d2 = torch.empty((N,), dtype=torch.float32, device=c.device)
            with _capi.guard(c):
                _capi.call("amc3d_crop_dist2", N, ptr(c), int(init_idx), ptr(d2), stream(c))
            crop_idx = torch.sort(d2, stable=True)[1][:voxel_max]
        elif not variable:
            if padding_choice is None:
                padding_choice = torch.randint(N, (voxel_max - N,), device=coord.device, generator=generator)
            crop_idx = torch.cat([torch.arange(N, device=coord.device), torch.as_tensor(padding_choice, device=coord.device).long()])
        if crop_idx is None:
            crop_idx = torch.arange(N, device=coord.device)
        if shuffle:
            if shuffle_perm is None:
                shuffle_perm = torch.randperm(crop_idx.numel(), device=coord.device, generator=generator)
            crop_idx = crop_idx[torch.as_tensor(shuffle_perm, device=coord.device).long()]
        coord = coord[crop_idx]
        feat = feat[crop_idx] if feat is not None else None
        label = label[crop_idx] if label is not None else None
    coord = coord - coord.min(0)[0]
    return (coord.float(), feat.float() if feat is not None else None, label.long() if label is not None else None)
