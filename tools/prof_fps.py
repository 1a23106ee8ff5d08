"""FPS timing (CUDA events).  `python tools/prof_fps.py N [M] [--chain]`:
default: 8 raw scenes of N points -> M = N/4 picks (level 1 of config 2 for N = 24000);
--chain: the PointNeXt chain 24000 -> 6000 -> 1500 -> 375 -> 93, where every level below the first runs on
a cloud that is already in FPS order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample


def timed(p, m):
    for _ in range(3):
        idx = furthest_point_sample(p, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    idx = furthest_point_sample(p, m)
    e1.record()
    torch.cuda.synchronize()
    return idx, e0.elapsed_time(e1)


args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 24000
if "--chain" in sys.argv:
    xyz, _ = scenes.batch_of_scenes(8, n, "surface")
    p = torch.from_numpy(xyz).cuda()
    tot = 0.0
    while p.shape[1] >= 4 * 93:
        m = p.shape[1] // 4
        idx, ms = timed(p, m)
        tot += ms
        print(f"fps n={p.shape[1]} m={m}: {ms:.4f} ms, {1e3 * ms / m:.4f} us/pick")
        p = torch.gather(p, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    print(f"chain total {tot:.4f} ms")
else:
    m = int(args[1]) if len(args) > 1 else n // 4
    xyz, _ = scenes.batch_of_scenes(8, n, "surface")
    p = torch.from_numpy(xyz).cuda()
    _, ms = timed(p, m)
    print(f"fps n={n} m={m}: {ms:.4f} ms, {1e3 * ms / m:.4f} us/pick")
