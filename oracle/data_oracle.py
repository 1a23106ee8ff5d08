"""CPU restatement (numpy) of openpoints/dataset/data_util.py: fnv_hash_vec (:92-105), ravel_hash_vec (:108-122),
voxelize (:125-143) and crop_pc (:137-174), with the reference's random draws turned into arguments and its
unstable argsort replaced by a stable one (the determinism contract of amcontrast3d_b200/data_util.py).

TEST INFRASTRUCTURE ONLY.  Pinned against outputs of the reference's own functions
(tests/golden/data_golden.npz, tests/golden/make_data_golden.py) by tests/test_oracle_data_golden.py.
"""
from __future__ import annotations

import numpy as np


def fnv_hash_vec(arr):
    arr = arr.copy().astype(np.uint64, copy=False)
    h = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= np.uint64(1099511628211)
        h = np.bitwise_xor(h, arr[:, j])
    return h


def ravel_hash_vec(arr):
    arr = arr.copy()
    arr -= arr.min(0)
    arr = arr.astype(np.uint64, copy=False)
    arr_max = arr.max(0).astype(np.uint64) + 1
    keys = np.zeros(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1] - 1):
        keys += arr[:, j]
        keys *= arr_max[j + 1]
    keys += arr[:, -1]
    return keys


def voxelize(coord, voxel_size=0.05, hash_type="fnv", mode=0, rand=None):
    discrete = np.floor(coord / np.array(voxel_size))
    key = ravel_hash_vec(discrete) if hash_type == "ravel" else fnv_hash_vec(discrete)
    idx_sort = np.argsort(key, kind="stable")
    key_sort = key[idx_sort]
    _, voxel_idx, count = np.unique(key_sort, return_counts=True, return_inverse=True)
    if mode == 0:
        start = np.cumsum(np.insert(count, 0, 0)[0:-1])
        return idx_sort[start + np.asarray(rand) % count]
    return idx_sort, voxel_idx, count


def crop_pc(coord, feat, label, split="train", voxel_size=0.04, voxel_max=None, downsample=True, variable=True,
            shuffle=True, rand=None, init_idx=None, shuffle_perm=None):
    coord = coord.copy()
    if voxel_size and downsample:
        coord -= coord.min(0)
        uniq = voxelize(coord, voxel_size, rand=rand)
        coord, feat, label = coord[uniq], feat[uniq] if feat is not None else None, label[uniq] if label is not None else None
    if voxel_max is not None:
        crop_idx = None
        N = len(label)
        if N >= voxel_max:
            if init_idx is None:
                init_idx = N // 2
            crop_idx = np.argsort(np.sum(np.square(coord - coord[init_idx]), 1), kind="stable")[:voxel_max]
        crop_idx = np.arange(coord.shape[0]) if crop_idx is None else crop_idx
        if shuffle:
            crop_idx = crop_idx[np.asarray(shuffle_perm)]
        coord, feat, label = coord[crop_idx], feat[crop_idx] if feat is not None else None, label[crop_idx] if label is not None else None
    coord -= coord.min(0)
    return coord.astype(np.float32), feat.astype(np.float32) if feat is not None else None, label.astype(np.int64) if label is not None else None
