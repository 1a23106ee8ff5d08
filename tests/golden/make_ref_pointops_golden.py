"""Run the REFERENCE's own CUDA kernels for the remaining pointops operators (oracle/_ref/libref_pointops.so,
compiled unmodified from /root/reference for sm_100) on seeded inputs on a B200 and store their outputs.

Run on the GPU box:   python tests/golden/make_ref_pointops_golden.py gpurun_out/ref_pointops_golden.npz
then copy the file to tests/golden/.  tests/test_oracle_ref_pointops_golden.py checks the oracle_pop_*
restatements of oracle/ops_oracle.c against it on CPU.  Inputs are regenerated from seeds by
golden_inputs() so only outputs are stored."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from amcontrast3d_b200 import scenes  # noqa: E402


def golden_inputs():
    rng = np.random.default_rng(4711)
    xyz, _ = scenes.batch_of_scenes(2, 2000, "surface", first_scene=70)
    xyz = np.ascontiguousarray(xyz.reshape(-1, 3))
    lattice = (rng.integers(0, 8, size=(3000, 3)).astype(np.float32) * 0.25)      # exact ties
    offset = np.array([1200, 1500, 4000], dtype=np.int32)                          # ragged segments
    new_offset = np.array([300, 340, 900], dtype=np.int32)
    lat_offset = np.array([1000, 3000], dtype=np.int32)
    lat_new_offset = np.array([250, 700], dtype=np.int32)
    n, ns, c, w_c, k, m = 600, 8, 24, 8, 3, 450
    feat = rng.standard_normal((n, c)).astype(np.float32)
    feat2 = rng.standard_normal((n, c)).astype(np.float32)
    nidx = rng.integers(0, n, size=(n, ns)).astype(np.int32)
    pos = rng.standard_normal((n, ns, c)).astype(np.float32)
    wgt = rng.standard_normal((n, ns, w_c)).astype(np.float32)
    g_nc = rng.standard_normal((n, c)).astype(np.float32)
    g_nsc = rng.standard_normal((n, ns, c)).astype(np.float32)
    src = rng.standard_normal((m, c)).astype(np.float32)                           # interpolation: (m,c) -> (n,c)
    iidx = rng.integers(0, m, size=(n, k)).astype(np.int32)
    iw = rng.random((n, k)).astype(np.float32)
    iw /= iw.sum(-1, keepdims=True)
    return dict(xyz=xyz, lattice=lattice, offset=offset, new_offset=new_offset, lat_offset=lat_offset,
                lat_new_offset=lat_new_offset, feat=feat, feat2=feat2, nidx=nidx, pos=pos, wgt=wgt, g_nc=g_nc,
                g_nsc=g_nsc, src=src, iidx=iidx, iw=iw.astype(np.float32))


def main(path):
    import torch
    from oracle import ref_kernels as rk
    inp = golden_inputs()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda")
    c = lambda x: x.cpu().numpy()
    out = {}
    xyz, off, noff = t(inp["xyz"]), t(inp["offset"]), t(inp["new_offset"])
    idx, tmp = rk.pop_furthestsampling(xyz, off, noff)
    out["fps/idx"], out["fps/tmp"] = c(idx), c(tmp)
    lat, loff, lnoff = t(inp["lattice"]), t(inp["lat_offset"]), t(inp["lat_new_offset"])
    idx, tmp = rk.pop_furthestsampling(lat, loff, lnoff)
    out["fps/lattice/idx"], out["fps/lattice/tmp"] = c(idx), c(tmp)
    q = xyz[out["fps/idx"].astype(np.int64)].contiguous()
    for r, ns in ((0.1, 16), (0.25, 8), (0.02, 4)):
        out[f"ballquery/{r}_{ns}"] = c(rk.pop_ballquery(r, ns, xyz, q, off, noff))
    out["ballquery/lattice"] = c(rk.pop_ballquery(0.5, 12, lat, lat, loff, loff))
    out["interpolation/fwd"] = c(rk.pop_interpolation_fwd(t(inp["src"]), t(inp["iidx"]), t(inp["iw"])))
    out["interpolation/bwd"] = c(rk.pop_interpolation_bwd(t(inp["g_nc"]), t(inp["iidx"]), t(inp["iw"]), inp["src"].shape[0]))
    out["subtraction/fwd"] = c(rk.pop_subtraction_fwd(t(inp["feat"]), t(inp["feat2"]), t(inp["nidx"])))
    g1, g2 = rk.pop_subtraction_bwd(t(inp["nidx"]), t(inp["g_nsc"]))
    out["subtraction/g1"], out["subtraction/g2"] = c(g1), c(g2)
    out["aggregation/fwd"] = c(rk.pop_aggregation_fwd(t(inp["feat"]), t(inp["pos"]), t(inp["wgt"]), t(inp["nidx"])))
    gi, gp, gw = rk.pop_aggregation_bwd(t(inp["feat"]), t(inp["pos"]), t(inp["wgt"]), t(inp["nidx"]), t(inp["g_nc"]))
    out["aggregation/gi"], out["aggregation/gp"], out["aggregation/gw"] = c(gi), c(gp), c(gw)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "gpurun_out", "ref_pointops_golden.npz"))
