import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample
for B, n in ((8, 24000), (2, 40000), (2, 64000), (8, 64000), (2, 100000), (2, 200000), (1, 400000)):
    xyz, _ = scenes.batch_of_scenes(B, n, "surface", first_scene=1)
    p = torch.from_numpy(xyz).cuda()
    m = n // 4
    furthest_point_sample(p, m); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); furthest_point_sample(p, m); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"B{B} n{n} -> {m}: {ms:.3f} ms  {1e3 * ms / m:.3f} us/pick", flush=True)
