"""us per pick of the cluster FPS at one scene size for the warps-per-CTA the dispatcher could choose (AMC3D_FPS_NW)"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample
for n in [int(a) for a in sys.argv[1:]] or [64000]:
    m = n // 4
    xyz, _ = scenes.batch_of_scenes(2, n, "surface", first_scene=1)
    p = torch.from_numpy(xyz).cuda()
    for _ in range(2): out = furthest_point_sample(p, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = furthest_point_sample(p, m); e1.record(); torch.cuda.synchronize()
    print(f"NW={os.environ.get('AMC3D_FPS_NW', 'auto')} n={n} m={m}: {e0.elapsed_time(e1) * 1e3 / m:.3f} us/pick  checksum {int(out.long().sum())}")
