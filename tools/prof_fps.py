"""FPS at the level-1 shape of config 2 (8 x 24000 -> 6000) — target for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
m = int(sys.argv[2]) if len(sys.argv) > 2 else n // 4
xyz, _ = scenes.batch_of_scenes(8, n, "surface")
p = torch.from_numpy(xyz).cuda()
for _ in range(3):
    idx = furthest_point_sample(p, m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
idx = furthest_point_sample(p, m)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"fps n={n} m={m}: {ms:.4f} ms, {1e3 * ms / m:.4f} us/round")
