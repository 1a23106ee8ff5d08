"""Shared helpers for the test-suite (inputs shared by the CPU oracle tests and the GPU parity tests)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from amcontrast3d_b200 import scenes  # noqa: E402
from oracle import ops_oracle  # noqa: E402


def args_ns(**kw):
    d = dict(nsample=16, ccbeta=0.04, cctype="Method2", temperature=0.3, supervisedCL="Method1", db="-m",
             margin="adaptive", mu=-1, nu=0.5, stages="up", stages_num=4, vis=False, w1=0.1, w2=0.9, w3=0.01)
    d.update(kw)
    return SimpleNamespace(**d)


# must stay identical to tests/golden/make_loss_golden.py CASES
GOLDEN_CASES = {
    "aa_default": dict(args=args_ns(), num_classes=13, ignore_index=None),
    "aa_k24": dict(args=args_ns(nsample=24), num_classes=13, ignore_index=None),
    "scannet": dict(args=args_ns(temperature=0.5, nu=0.6), num_classes=20, ignore_index=-100),
    "m3_plus_cl2": dict(args=args_ns(cctype="Method3", db="+m", supervisedCL="Method2", temperature=None),
                        num_classes=13, ignore_index=None),
    "m1_const": dict(args=args_ns(cctype="Method1", margin="constant", db="none"), num_classes=13,
                     ignore_index=None),
}


def build_hierarchy(seed=7, batch=2, n0=1024, dims=(32, 64, 128, 256), num_classes=13, ignore_fraction=0.0,
                    kind="surface"):
    """Same construction as make_loss_golden.build_inputs: flattened 4-stage hierarchy, FPS /4 per scene."""
    xyz, lab = scenes.batch_of_scenes(batch, n0, kind, first_scene=seed, num_classes=num_classes,
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


def stage_list_of(p_list, f_list, device="cpu", requires_grad=True):
    down = []
    for p, f in zip(p_list, f_list):
        ft = torch.from_numpy(np.array(f, copy=True)).to(device)
        ft.requires_grad_(requires_grad)
        down.append({"p_out": torch.from_numpy(np.array(p, copy=True)).to(device), "f_out": ft,
                     "offset": torch.tensor([p.shape[0]], dtype=torch.int32, device=device)})
    return {"inputs": None, "down": down, "up": down}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
