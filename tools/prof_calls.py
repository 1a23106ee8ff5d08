"""Per-call device times of one path-replay step (CUDA events around every C-ABI call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200 import _capi
from amcontrast3d_b200.replay import PathReplay

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
r = PathReplay(batch=8, n_points=24000, k=k)
for _ in range(3):
    r.step()
torch.cuda.synchronize()
_capi.PROFILE = []
r.step()
torch.cuda.synchronize()
prof, _capi.PROFILE = _capi.PROFILE, None
tot = 0.0
for name, e0, e1, a in prof:
    ms = e0.elapsed_time(e1)
    tot += ms
    print(f"{name:34s} {ms:9.4f} ms  args={tuple(int(x) if isinstance(x, int) else round(float(x), 3) if isinstance(x, float) else x for x in a[:5])}")
print("total", tot)
