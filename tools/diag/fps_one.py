import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
xyz, _ = scenes.batch_of_scenes(2, n, "surface", first_scene=1)
p = torch.from_numpy(xyz).cuda()
for _ in range(2):
    furthest_point_sample(p, m)
torch.cuda.synchronize()
