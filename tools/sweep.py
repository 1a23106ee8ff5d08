"""BASELINE configs 3 and 5 as measurements (CUDA events, eager calls):
  * config 3: AMContrast3D++ (MM) replay — ScanNet-shaped B=2 x 64000 points, 20 classes + ignored labels,
    DualMasks refinement (K=8) before the loss;
  * config 5: scaling sweep over points per scene (B=1), k and embedding width for the three pieces of the
    path separately: self-kNN, grouping gather + scatter-add (C=64, nsample=32), AM loss forward+backward.
Prints markdown tables (committed under profiles/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from amcontrast3d_b200 import _amloss, scenes
from amcontrast3d_b200.AMContrast3D import ContrastHead
from amcontrast3d_b200.layers import ball_query, furthest_point_sample, grouping_operation
from amcontrast3d_b200.replay import PathReplay, aa_args


def t(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("## config 3 — AMContrast3D++ replay (B=2 x 64000, 20 classes + ignore, DualMasks K=8, k=12)\n")
r = PathReplay(batch=2, n_points=64000, k=12, num_classes=20, ignore_index=-100, refine=True, refine_k=8)
ms = t(lambda: r.step(), reps=5, warm=3)
print(f"eager step {ms:.2f} ms -> {2 * 64000 / ms * 1e3 / 1e6:.2f} M scene-points/s")
r.capture(warmup=1)
ms = t(lambda: r.step_graph(), reps=10, warm=3)
print(f"graph replay {ms:.2f} ms -> {2 * 64000 / ms * 1e3 / 1e6:.2f} M scene-points/s\n")
del r
torch.cuda.empty_cache()

print("## config 5 — sweep, one scene per call\n")
print("| points | FPS n->n/4 ms | kNN k=16 ms | kNN k=32 ms | group fwd ms (GB/s) | group bwd ms (GB/s) | loss f+b D=32 | D=64 | D=128 | D=256 (ms, k=16) |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for n in (10000, 24000, 50000, 100000, 200000):
    xyz, lab = scenes.batch_of_scenes(1, n, "surface", first_scene=3)
    p = torch.from_numpy(xyz).cuda()
    flat = p.reshape(-1, 3).contiguous()
    off = torch.tensor([n], dtype=torch.int32, device="cuda")
    target = torch.from_numpy(lab.reshape(-1)).cuda()
    fps_ms = t(lambda: furthest_point_sample(p, n // 4))
    k16 = t(lambda: _amloss.knn_raw(16, flat, flat, off, off))
    k32 = t(lambda: _amloss.knn_raw(32, flat, flat, off, off))
    q = torch.gather(p, 1, furthest_point_sample(p, n // 4).long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx = ball_query(0.1, 32, p, q)
    f = torch.randn(1, 64, n, device="cuda", requires_grad=True)
    out = grouping_operation(f, idx)
    g = torch.randn_like(out)
    gf = t(lambda: grouping_operation(f, idx))

    def bwd():
        f.grad = None
        o = grouping_operation(f, idx)
        o.backward(g)
    gb = t(bwd) - gf
    by = 4.0 * 64 * (n // 4) * 32
    row = [f"{n}", f"{fps_ms:.3f}", f"{k16:.3f}", f"{k32:.3f}", f"{gf:.3f} ({by / gf / 1e6:.0f})", f"{gb:.3f} ({by / gb / 1e6:.0f})"]
    head = ContrastHead()
    args = aa_args(16, stages_num=1)
    for D in (32, 64, 128, 256):
        feat = torch.randn(n, D, device="cuda", requires_grad=True)
        st = {"p_out": flat, "f_out": feat, "offset": off}
        sl = {"inputs": None, "down": [st], "up": [st]}

        def loss_fb():
            feat.grad = None
            l, _, _ = head(None, target, sl, 13, None, args)
            l.backward()
        row.append(f"{t(loss_fb):.3f}")
    print("| " + " | ".join(row) + " |")
    del f, out, g
    torch.cuda.empty_cache()
