"""amcontrast3d_b200.data_util (voxelize / crop_pc on CUDA tensors) against the numpy restatement
(oracle/data_oracle.py, same draws, same stable order -> bit-identical) and against the reference's own outputs
(tests/golden/data_golden.npz: keys order, voxel_idx, counts identical; members as sets)."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import REPO
from oracle import data_oracle as do

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
PATH = os.path.join(REPO, "tests", "golden", "data_golden.npz")
pytestmark = pytest.mark.gpu


def _inputs():
    from make_data_golden import inputs
    return inputs()


@pytest.mark.parametrize("hash_type", ["fnv", "ravel"])
def test_voxelize_val_mode(hash_type):
    from amcontrast3d_b200 import data_util as du
    g = np.load(PATH)
    coord, _, _ = _inputs()
    c0 = coord - coord.min(0)
    idx_sort, voxel_idx, count = du.voxelize(torch.from_numpy(c0).cuda(), 0.04, hash_type, mode=1)
    r_idx, r_vox, r_cnt = do.voxelize(c0, 0.04, hash_type, mode=1)
    assert np.array_equal(idx_sort.cpu().numpy(), r_idx)                       # the restatement: bit-identical
    assert np.array_equal(voxel_idx.cpu().numpy(), r_vox) and np.array_equal(count.cpu().numpy(), r_cnt)
    assert np.array_equal(count.cpu().numpy(), g[f"vox_{hash_type}/count"])   # the reference itself
    assert np.array_equal(voxel_idx.cpu().numpy(), g[f"vox_{hash_type}/voxel_idx"])


def test_voxel_keys_are_the_reference_hash():
    from amcontrast3d_b200 import data_util as du
    rng = np.random.default_rng(0)
    coord = (rng.random((5000, 3)) * 7).astype(np.float32)
    key, cells = du._keys(torch.from_numpy(coord).cuda(), 0.05, want_cells=True)
    disc = np.floor(coord / np.array(0.05))
    assert np.array_equal(cells.cpu().numpy(), disc.astype(np.int64))
    assert np.array_equal(key.cpu().numpy().view(np.uint64), do.fnv_hash_vec(disc))
    assert np.array_equal(du.fnv_hash_vec(cells).cpu().numpy().view(np.uint64), do.fnv_hash_vec(disc))
    assert np.array_equal(du.ravel_hash_vec(cells).cpu().numpy().view(np.uint64), do.ravel_hash_vec(disc))


def test_voxelize_train_mode_and_crop_pc_match_the_restatement():
    from amcontrast3d_b200 import data_util as du
    g = np.load(PATH)
    coord, feat, label = _inputs()
    c0 = coord - coord.min(0)
    uniq = du.voxelize(torch.from_numpy(c0).cuda(), 0.04, rand=torch.from_numpy(g["vox_train/rand"]))
    assert np.array_equal(uniq.cpu().numpy(), do.voxelize(c0, 0.04, rand=g["vox_train/rand"]))
    cnt = g["vox_fnv/count"]
    assert np.array_equal(uniq.cpu().numpy()[cnt == 1], g["vox_train/uniq"][cnt == 1])       # the reference's picks
    for tag, shuffle in (("crop", False), ("crop_shuf", True)):
        perm = g[f"{tag}/perm"] if shuffle else None
        c, f, l = du.crop_pc(torch.from_numpy(coord).cuda(), torch.from_numpy(feat).cuda(), torch.from_numpy(label).cuda(),
                             split="val", voxel_size=0.04, voxel_max=3000, shuffle=shuffle,
                             rand=torch.from_numpy(g[f"{tag}/rand"]),
                             shuffle_perm=torch.from_numpy(perm) if shuffle else None)
        rc, rf, rl = do.crop_pc(coord, feat, label, split="val", voxel_size=0.04, voxel_max=3000, shuffle=shuffle,
                                rand=g[f"{tag}/rand"], shuffle_perm=perm)
        assert c.dtype == torch.float32 and l.dtype == torch.int64
        assert np.array_equal(c.cpu().numpy(), rc) and np.array_equal(f.cpu().numpy(), rf) and np.array_equal(l.cpu().numpy(), rl)


def test_generator_driven_randomness_is_reproducible():
    from amcontrast3d_b200 import data_util as du
    coord, feat, label = _inputs()
    args = (torch.from_numpy(coord).cuda(), torch.from_numpy(feat).cuda(), torch.from_numpy(label).cuda())
    outs = []
    for _ in range(2):
        gen = torch.Generator(device="cuda").manual_seed(123)
        outs.append(du.crop_pc(*args, split="train", voxel_size=0.04, voxel_max=3000, generator=gen))
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert outs[0][0].shape == (3000, 3) and float(outs[0][0].min()) == 0.0
