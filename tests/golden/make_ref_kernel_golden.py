"""Run the REFERENCE's own CUDA kernels (oracle/_ref/libref_kernels.so, compiled unmodified from
/root/reference for sm_100) on seeded inputs on a B200 and store their outputs.

Run on the GPU box:   python tests/golden/make_ref_kernel_golden.py gpurun_out/ref_kernels_golden.npz
then copy the file to tests/golden/.  tests/test_oracle_ref_golden.py checks oracle/ops_oracle.c
against it on CPU — this is what pins the C restatement to the real reference kernels.
Inputs are regenerated from seeds by golden_inputs() so only outputs are stored."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from amcontrast3d_b200 import scenes  # noqa: E402


def golden_inputs():
    rng = np.random.default_rng(2024)
    xyz, _ = scenes.batch_of_scenes(2, 3000, "surface", first_scene=40)
    lattice = rng.integers(0, 10, size=(2, 1500, 3)).astype(np.float32) * 0.25     # exact ties
    feats = rng.standard_normal((2, 24, 3000)).astype(np.float32)
    gidx = rng.integers(0, 3000, size=(2, 200, 8)).astype(np.int32)
    go = rng.standard_normal((2, 24, 200, 8)).astype(np.float32)
    w = rng.random((2, 3000, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    segs = np.array([1000, 1007, 3000], dtype=np.int32)     # ragged segments, one shorter than k
    return dict(xyz=xyz, lattice=lattice, feats=feats, gidx=gidx, go=go, w=w.astype(np.float32), segs=segs)


def main(path):
    import torch
    from oracle import ref_kernels as rk
    inp = golden_inputs()
    dev = "cuda"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = {}
    xyz = t(inp["xyz"])
    for m in (750, 100):
        idx, temp = rk.fps(xyz, m)
        out[f"fps/{m}/idx"], out[f"fps/{m}/temp"] = idx.cpu().numpy(), temp.cpu().numpy()
    idx, temp = rk.fps(t(inp["lattice"]), 400)
    out["fps/lattice/idx"], out["fps/lattice/temp"] = idx.cpu().numpy(), temp.cpu().numpy()
    q = xyz[:, :750].contiguous()
    for r, ns in ((0.1, 32), (0.2, 16)):
        out[f"ball_query/{r}_{ns}"] = rk.ball_query(r, ns, xyz, q).cpu().numpy()
    d2, i3 = rk.three_nn(xyz, q)
    out["three_nn/dist2"], out["three_nn/idx"] = d2.cpu().numpy(), i3.cpu().numpy()
    feats = t(inp["feats"])
    out["three_interpolate/out"] = rk.three_interpolate(feats[:, :, :750].contiguous(), i3, t(inp["w"])).cpu().numpy()
    out["group_points/out"] = rk.group_points(feats, t(inp["gidx"])).cpu().numpy()
    out["group_points_grad/out"] = rk.group_points_grad(t(inp["go"]), t(inp["gidx"]), 3000).cpu().numpy()
    flat = xyz.reshape(-1, 3).contiguous()
    o1 = t(np.array([6000], dtype=np.int32))
    for k in (4, 16, 24, 64):
        i, d = rk.knnquery(k, flat, flat, o1, o1)
        out[f"knn/{k}/idx"], out[f"knn/{k}/dist2"] = i.cpu().numpy(), d.cpu().numpy()
    so = np.concatenate([inp["segs"], [6000]]).astype(np.int32)
    i, d = rk.knnquery(12, flat, flat, t(so), t(so))
    out["knn/segs/idx"], out["knn/segs/dist2"] = i.cpu().numpy(), d.cpu().numpy()
    lat = t(inp["lattice"].reshape(-1, 3))
    ol = t(np.array([3000], dtype=np.int32))
    i, d = rk.knnquery(8, lat, lat, ol, ol)                      # heap behaviour under exact ties
    out["knn/lattice/idx"], out["knn/lattice/dist2"] = i.cpu().numpy(), d.cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ref_kernels_golden.npz")
