"""Times of the culled searches at the config-2 shapes: kNN (loss stages + label votes), three_nn, ball_query."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from amcontrast3d_b200 import _amloss, scenes
from amcontrast3d_b200.layers import ball_query, furthest_point_sample, three_nn


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
p = [torch.from_numpy(xyz).cuda()]
for n in (6000, 1500, 375):
    i = furthest_point_sample(p[-1], n).long()
    p.append(torch.gather(p[-1], 1, i.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
flat = [x.reshape(-1, 3).contiguous() for x in p]
off = [torch.tensor([f.shape[0]], dtype=torch.int32, device="cuda") for f in flat]
tot = 0.0
for s in range(4):
    ms = t(lambda: _amloss.knn_raw(16, flat[s], flat[s], off[s], off[s]))
    tot += ms
    print(f"knn self stage {s} m={flat[s].shape[0]:6d} k=16: {ms:.4f} ms")
for s, kr in ((1, 4), (2, 16), (3, 64)):
    ms = t(lambda: _amloss.knn_raw(kr, flat[0], flat[s], off[0], off[s]))
    tot += ms
    print(f"knn vote  stage {s} m={flat[s].shape[0]:6d} k={kr:2d}: {ms:.4f} ms")
print(f"knn total {tot:.4f} ms")
tot = 0.0
for l in (1, 2):
    ms = t(lambda: three_nn(p[l - 1], p[l]))
    tot += ms
    print(f"three_nn {p[l - 1].shape[1]} <- {p[l].shape[1]}: {ms:.4f} ms")
print(f"three_nn total {tot:.4f} ms")
tot = 0.0
for (ls, lq, r, cnt) in ((0, 1, 0.1, 1), (1, 1, 0.2, 3), (1, 2, 0.2, 1), (2, 2, 0.4, 6)):
    ms = t(lambda: ball_query(r, 32, p[ls], p[lq]))
    tot += ms * cnt
    print(f"ball_query N={p[ls].shape[1]} M={p[lq].shape[1]} r={r}: {ms:.4f} ms x{cnt}")
print(f"ball total {tot:.4f} ms")
