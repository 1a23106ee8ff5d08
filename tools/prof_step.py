"""Run N eager path-replay steps of config 2 (target for `ncu -k regex:... -s <skip> -c <count>`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200.replay import PathReplay

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
r = PathReplay(batch=8, n_points=24000, k=16)
for _ in range(steps):
    r.step()
torch.cuda.synchronize()
print("ok", steps)
