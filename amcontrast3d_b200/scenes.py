"""Deterministic synthetic scenes shaped like the reference's training crops.

There is no dataset on the box, so the benchmark and the tests use these generators.  Shapes
follow the reference's data pipeline: S3DIS crops of `voxel_max` = 24 000 points taken as the
N nearest points to a random centre, shifted to the origin and shuffled
(openpoints/dataset/data_util.py:149-173, cfgs/s3dis/default.yaml:10); ScanNet crops of
64 000 points with ~5 % unlabelled (-100) points (cfgs/scannet/default.yaml:9).

numpy only — no torch, no CUDA: both the product bench and the CPU oracle consume the same
arrays.  Seeds: 1234 + 1000*rank + scene_id.
"""
from __future__ import annotations

import numpy as np


def scene_seed(scene_id: int, rank: int = 0) -> int:
    return 1234 + 1000 * rank + scene_id


def _lattice_plane(rng, origin, u, v, lu, lv, step):
    nu, nv = max(int(lu / step), 1), max(int(lv / step), 1)
    gu, gv = np.meshgrid(np.arange(nu) * step, np.arange(nv) * step, indexing="ij")
    return origin[None, :] + gu.reshape(-1, 1) * u[None, :] + gv.reshape(-1, 1) * v[None, :]


def surface_scene(n_points: int = 24000, seed: int = 1234, num_classes: int = 13, step: float = 0.04,
                  ignore_fraction: float = 0.0, ignore_index: int = -100):
    """S3DIS-shaped room: floor, ceiling, 4 walls and 14 axis-aligned boxes sampled on a `step`
    lattice with +-0.015 m jitter per coordinate (keeps points distinct => tie-free kNN/FPS).
    Labels: ceiling 0, floor 1, wall 2, box 3 + (id mod (num_classes-3)).
    Returns xyz (n,3) float32 >= 0 and labels (n,) int64."""
    rng = np.random.default_rng(seed)
    W, L, H = rng.uniform(5, 9), rng.uniform(5, 9), 3.0
    ex, ey, ez = np.eye(3)
    pts, lab = [], []

    def add(p, l):
        pts.append(p)
        lab.append(np.full(len(p), l, dtype=np.int64))

    add(_lattice_plane(rng, np.array([0, 0, H]), ex, ey, W, L, step), 0)
    add(_lattice_plane(rng, np.array([0, 0, 0.0]), ex, ey, W, L, step), 1)
    add(_lattice_plane(rng, np.array([0, 0, 0.0]), ex, ez, W, H, step), 2)
    add(_lattice_plane(rng, np.array([0, L, 0.0]), ex, ez, W, H, step), 2)
    add(_lattice_plane(rng, np.array([0, 0, 0.0]), ey, ez, L, H, step), 2)
    add(_lattice_plane(rng, np.array([W, 0, 0.0]), ey, ez, L, H, step), 2)
    nbox_cls = max(num_classes - 3, 1)
    for bid in range(14):
        sx, sy, sz = rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.8)
        ox, oy = rng.uniform(0.1, W - sx - 0.1), rng.uniform(0.1, L - sy - 0.1)
        o = np.array([ox, oy, 0.0])
        l = 3 + (bid % nbox_cls)
        add(_lattice_plane(rng, o + np.array([0, 0, sz]), ex, ey, sx, sy, step), l)      # top
        add(_lattice_plane(rng, o, ex, ez, sx, sz, step), l)
        add(_lattice_plane(rng, o + np.array([0, sy, 0]), ex, ez, sx, sz, step), l)
        add(_lattice_plane(rng, o, ey, ez, sy, sz, step), l)
        add(_lattice_plane(rng, o + np.array([sx, 0, 0]), ey, ez, sy, sz, step), l)
    xyz = np.concatenate(pts)
    labels = np.concatenate(lab)
    xyz = xyz + rng.uniform(-0.015, 0.015, size=xyz.shape)
    # crop: the n nearest points to a random point of the cloud (data_util.py:158-160)
    if len(xyz) < n_points:
        reps = int(np.ceil(n_points / len(xyz)))
        extra = rng.uniform(-0.004, 0.004, size=(reps * len(xyz), 3))
        xyz = np.tile(xyz, (reps, 1)) + extra
        labels = np.tile(labels, reps)
    centre = xyz[rng.integers(len(xyz))]
    order = np.argsort(((xyz - centre) ** 2).sum(1), kind="stable")[:n_points]
    xyz, labels = xyz[order], labels[order]
    xyz = xyz - xyz.min(0)                      # data_util.py:173
    perm = rng.permutation(n_points)            # data_util.py:169-171
    xyz, labels = xyz[perm].astype(np.float32), labels[perm]
    if ignore_fraction > 0:
        labels = labels.copy()
        labels[rng.random(n_points) < ignore_fraction] = ignore_index
    return np.ascontiguousarray(xyz), labels


def volume_scene(n_points: int = 24000, seed: int = 1234, num_classes: int = 13):
    """Stress scene: uniform points in 4 x 4 x 3 m, label = nearest of 52 random seeds mod
    num_classes (about half of the points are boundary points at k = 16)."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(0, 1, size=(n_points, 3)) * np.array([4.0, 4.0, 3.0])
    seeds = rng.uniform(0, 1, size=(52, 3)) * np.array([4.0, 4.0, 3.0])
    d = ((xyz[:, None, :] - seeds[None, :, :]) ** 2).sum(-1)
    labels = (d.argmin(1) % num_classes).astype(np.int64)
    return np.ascontiguousarray(xyz.astype(np.float32)), labels


def batch_of_scenes(batch: int, n_points: int, kind: str = "surface", rank: int = 0, first_scene: int = 0,
                    num_classes: int = 13, ignore_fraction: float = 0.0):
    """(B,N,3) float32 xyz and (B,N) int64 labels for scenes first_scene .. first_scene+batch-1."""
    xs, ls = [], []
    for s in range(first_scene, first_scene + batch):
        if kind == "surface":
            x, l = surface_scene(n_points, scene_seed(s, rank), num_classes, ignore_fraction=ignore_fraction)
        elif kind == "volume":
            x, l = volume_scene(n_points, scene_seed(s, rank), num_classes)
        else:
            raise ValueError(kind)
        xs.append(x)
        ls.append(l)
    return np.stack(xs), np.stack(ls)
