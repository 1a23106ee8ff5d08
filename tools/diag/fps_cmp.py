import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample
for B, n, m in ((1, 230000, 3000), (1, 400000, 3000), (2, 1000000, 2000)):
    xyz, _ = scenes.batch_of_scenes(B, n, "surface", first_scene=1)
    p = torch.from_numpy(xyz).cuda()
    furthest_point_sample(p, 64); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); furthest_point_sample(p, m); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B{B} n{n} -> {m}: {ms:.3f} ms  {1e3 * ms / m:.3f} us/pick", flush=True)
