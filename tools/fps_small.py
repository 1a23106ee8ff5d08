import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import furthest_point_sample
from oracle import ops_oracle as oo
n = int(sys.argv[1]); m = int(sys.argv[2]); b = int(sys.argv[3])
xyz, _ = scenes.batch_of_scenes(b, n, "surface")
idx = furthest_point_sample(torch.from_numpy(xyz).cuda(), m)
torch.cuda.synchronize()
ridx, _ = oo.fps(xyz, m)
print("match", np.array_equal(idx.cpu().numpy(), ridx))
