"""Generate tests/golden/data_golden.npz by running the REFERENCE's own voxelize / crop_pc
(openpoints/dataset/data_util.py:92-174) on seeded inputs.  Build container only (needs /root/reference):

    python tests/golden/make_data_golden.py

The reference draws from numpy's global random stream; the script seeds it and records the draws it made
(by replaying the same seed), so that the restatement and the GPU implementation can be given the same ones.
`np.long` (removed from numpy) is restored for the duration of the call; `h5py` is stubbed."""
import os
import sys
import types

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
from amcontrast3d_b200 import scenes  # noqa: E402


def inputs(seed=3, n=12000):
    xyz, lab = scenes.batch_of_scenes(1, n, "surface", first_scene=seed)
    rng = np.random.default_rng(seed)
    feat = rng.random((n, 3), dtype=np.float32)
    return xyz[0].astype(np.float32), feat, lab[0].astype(np.int64)


def main():
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if not hasattr(np, "long"):
        np.long = np.int64
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_data_util", "/root/reference/openpoints/dataset/data_util.py")
    du = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(du)
    out = {}
    coord, feat, label = inputs()
    c0 = coord - coord.min(0)
    # voxelize, val mode: (idx_sort, voxel_idx, count)
    for hash_type in ("fnv", "ravel"):
        idx_sort, voxel_idx, count = du.voxelize(c0.copy(), 0.04, hash_type, mode=1)
        out[f"vox_{hash_type}/idx_sort"], out[f"vox_{hash_type}/voxel_idx"], out[f"vox_{hash_type}/count"] = idx_sort, voxel_idx, count
    # voxelize, train mode, with the global stream seeded; the draw is reproduced from the same seed
    np.random.seed(11)
    uniq = du.voxelize(c0.copy(), 0.04)
    count = out["vox_fnv/count"]
    np.random.seed(11)
    out["vox_train/rand"] = np.random.randint(0, count.max(), count.size)
    out["vox_train/uniq"] = uniq
    # crop_pc, val split (init = N // 2), no shuffle and with shuffle
    np.random.seed(5)
    c, f, l = du.crop_pc(coord.copy(), feat.copy(), label.copy(), split="val", voxel_size=0.04, voxel_max=3000, shuffle=False)
    np.random.seed(5)
    nvox = count.size
    out["crop/rand"] = np.random.randint(0, count.max(), nvox)
    out["crop/coord"], out["crop/feat"], out["crop/label"] = c, f, l
    np.random.seed(7)
    c, f, l = du.crop_pc(coord.copy(), feat.copy(), label.copy(), split="val", voxel_size=0.04, voxel_max=3000, shuffle=True)
    np.random.seed(7)
    out["crop_shuf/rand"] = np.random.randint(0, count.max(), nvox)
    out["crop_shuf/perm"] = np.random.permutation(np.arange(3000))
    out["crop_shuf/coord"], out["crop_shuf/feat"], out["crop_shuf/label"] = c, f, l
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
