"""Generate tests/golden/loss_golden.npz by running the REFERENCE's own Python modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_loss_golden.py

The reference's loss path (openpoints/AMContrast3D/{MarginContrast,MaskedRefine,metrics}.py,
AEF/*) is imported unmodified from /root/reference and executed on CPU.  It has hard-coded
`.cuda()` calls and imports the compiled `pointops_cuda` extension, so (as in SURVEY.md §8c,
tier O2) this script provides:
  * sys.modules stubs for third-party packages that are absent here and irrelevant to the path,
  * a fake `pointops_cuda.knnquery_cuda` backed by oracle/ops_oracle.c (the literal restatement
    of the reference heap kernel), and an empty `pointnet2_batch_cuda`,
  * `Tensor.cuda()` -> identity, `torch.cuda.{Int,Float}Tensor` -> CPU constructors.
Nothing of the reference is copied: only its outputs on seeded inputs are stored.  The
fixtures pin oracle/loss_oracle.py (tests/test_oracle_golden.py) and, through it, the kernels.
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import ops_oracle  # noqa: E402
from amcontrast3d_b200 import scenes  # noqa: E402


def install_stubs():
    for name in ["wandb", "shortuuid", "termcolor", "h5py", "easydict", "multimethod", "torch_scatter",
                 "pyvista", "deepspeed", "torcheval", "torcheval.metrics"]:
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    sys.modules["easydict"].EasyDict = dict
    sys.modules["multimethod"].multimethod = lambda f: f
    sys.modules["torcheval.metrics"].R2Score = object

    pops = types.ModuleType("pointops_cuda")

    def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        i, d = ops_oracle.knnquery(int(nsample), xyz.numpy(), new_xyz.numpy(), offset.numpy(), new_offset.numpy())
        idx.copy_(torch.from_numpy(i))
        dist2.copy_(torch.from_numpy(d))

    pops.knnquery_cuda = knnquery_cuda
    sys.modules["pointops_cuda"] = pops
    sys.modules["pointnet2_batch_cuda"] = types.ModuleType("pointnet2_batch_cuda")

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.IntTensor = lambda *a, **k: torch.zeros(*a, dtype=torch.int32)
    torch.cuda.FloatTensor = lambda *a, **k: torch.zeros(*a, dtype=torch.float32)


class Args(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def base_args(**kw):
    a = Args(nsample=16, ccbeta=0.04, cctype="Method2", temperature=0.3, supervisedCL="Method1", db="-m",
             margin="adaptive", mu=-1, nu=0.5, stages="up", stages_num=4, vis=False, w1=0.1, w2=0.9, w3=0.01)
    a.update(kw)
    return a


CASES = {
    # shipped S3DIS setting at k=16 (cfgs/s3dis/AMContrast3D-AA.yaml:6-30)
    "aa_default": dict(args=base_args(), num_classes=13, ignore_index=None),
    # shipped nsample
    "aa_k24": dict(args=base_args(nsample=24), num_classes=13, ignore_index=None),
    # ScanNet setting: 20 classes + ignore_index -> class 20, T=0.5, nu=0.6
    "scannet": dict(args=base_args(temperature=0.5, nu=0.6), num_classes=20, ignore_index=-100),
    # the other enumerators of the flag surface
    "m3_plus_cl2": dict(args=base_args(cctype="Method3", db="+m", supervisedCL="Method2", temperature=None),
                        num_classes=13, ignore_index=None),
    "m1_const": dict(args=base_args(cctype="Method1", margin="constant", db="none"), num_classes=13,
                     ignore_index=None),
}


def build_inputs(seed=7, batch=2, n0=1024, dims=(32, 64, 128, 256), num_classes=13, ignore_fraction=0.0):
    """Flattened 4-stage hierarchy: FPS /4 per scene (oracle restatement of the reference FPS)."""
    xyz, lab = scenes.batch_of_scenes(batch, n0, "surface", first_scene=seed, num_classes=num_classes,
                                      ignore_fraction=ignore_fraction)
    rng = np.random.default_rng(seed)
    p_list, f_list = [], []
    cur = xyz
    for s, d in enumerate(dims):
        if s > 0:
            idx, _ = ops_oracle.fps(cur, cur.shape[1] // 4)
            cur = np.take_along_axis(cur, idx[:, :, None].astype(np.int64), axis=1)
        p_list.append(np.ascontiguousarray(cur.reshape(-1, 3)))
        f_list.append(rng.standard_normal((p_list[-1].shape[0], d)).astype(np.float32))
    return p_list, f_list, lab.reshape(-1)


def stage_list_of(p_list, f_list, requires_grad=True):
    down = []
    for p, f in zip(p_list, f_list):
        ft = torch.from_numpy(f.copy())
        ft.requires_grad_(requires_grad)
        down.append({"p_out": torch.from_numpy(p.copy()), "f_out": ft,
                     "offset": torch.IntTensor([p.shape[0]])})
    return {"inputs": None, "down": down, "up": down}


def main():
    install_stubs()
    import openpoints.models  # noqa: F401  (must come first: circular import otherwise)
    from openpoints.AMContrast3D.MarginContrast import AmbiguityHead, ContrastHead
    from openpoints.AMContrast3D.MaskedRefine import RefinementMethod
    from openpoints.AMContrast3D.metrics import posmask_searching
    from openpoints.AMContrast3D.AEF.utils import get_subscene_label_CBL

    out = {}
    for name, case in CASES.items():
        ncls, ign = case["num_classes"], case["ignore_index"]
        p_list, f_list, target = build_inputs(num_classes=ncls, ignore_fraction=0.05 if ign is not None else 0.0)
        if name in ("aa_default", "scannet"):
            for s in range(4):
                out[f"{name}/p{s}"] = p_list[s]
                out[f"{name}/f{s}"] = f_list[s]
            out[f"{name}/target"] = target
        sl = stage_list_of(p_list, f_list)
        tgt = torch.from_numpy(target)
        head = ContrastHead()
        loss, a_cat, a_list = head(None, tgt, sl, ncls, ign, Args(case["args"]))
        loss.backward()
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/a_cat"] = a_cat.detach().numpy()
        for s in range(4):
            g = sl["up"][s]["f_out"].grad.numpy()
            if name in ("aa_default", "m3_plus_cl2"):
                out[f"{name}/grad{s}"] = g
            out[f"{name}/grad{s}_norm"] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
            out[f"{name}/grad{s}_rows"] = g[:: max(1, g.shape[0] // 16)][:16].copy()
        print(name, "loss", loss.item(), "sel", [int(((a > 0) & (a <= 1)).sum()) for a in a_list])

    # AmbiguityHead, soft stage labels, posmask_searching on the default inputs
    case = CASES["aa_default"]
    p_list, f_list, target = build_inputs()
    sl = stage_list_of(p_list, f_list, requires_grad=False)
    tgt = torch.from_numpy(target)
    a_list = AmbiguityHead()(tgt, sl, 13, None, Args(case["args"]))
    out["ambiguity_head/a_cat"] = torch.cat(a_list).numpy()
    nstride = torch.tensor([4, 4, 4, 4])
    for s in range(4):
        out[f"labels/stage{s}"] = get_subscene_label_CBL("up", s, sl, tgt, nstride, 13, None).numpy()
    pm, nidx = posmask_searching(torch.from_numpy(p_list[0]), tgt, 16, 13, None)
    out["posmask_searching/posmask"] = pm.numpy()
    out["posmask_searching/nidx"] = nidx.numpy()

    # RefinementMethod.DualMasks, ScanNet-MM-like settings (cfgs/scannet/AMContrast3D-MM.yaml:39-52)
    rng = np.random.default_rng(11)
    B, n, D, K = 2, 512, 32, 8
    p = torch.from_numpy(np.ascontiguousarray(p_list[0].reshape(2, -1, 3)[:, :n]))
    for fusion, thr, thr_max, gamma in (("MIN", 0.9, 1.0, 0.4), ("MIN", 0.5, 0.8, 1.0), ("MIN_ALL0", 0.9, 1.0, 0.4)):
        f = torch.from_numpy(rng.standard_normal((B, D, n)).astype(np.float32)).requires_grad_(True)
        a = torch.from_numpy(rng.random((B, 1, n)).astype(np.float32))
        if fusion == "MIN_ALL0":
            a = torch.where(a < 0.3, torch.zeros_like(a), a)
        w = torch.from_numpy(rng.standard_normal((B, D, n)).astype(np.float32))
        ref = RefinementMethod(sl, p, f, a, -1, B, K, fusion, thr_max, thr, gamma)
        feat, rate = ref.DualMasks()
        (feat * w).sum().backward()
        tag = f"refine/{fusion}_{thr}_{gamma}"
        out[f"{tag}/p"] = p.numpy()
        out[f"{tag}/f"] = f.detach().numpy()
        out[f"{tag}/a"] = a.numpy()
        out[f"{tag}/w"] = w.numpy()
        out[f"{tag}/out"] = feat.detach().numpy()
        out[f"{tag}/grad"] = f.grad.numpy()
        out[f"{tag}/rate"] = np.float64(rate)
        print(tag, "rate", rate)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
