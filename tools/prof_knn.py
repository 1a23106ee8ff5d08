"""Run the stage-0 self-kNN (B*N = 192000 points, one segment) a few times — target for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from amcontrast3d_b200 import _amloss, scenes

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = int(sys.argv[2]) if len(sys.argv) > 2 else 192000
xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
flat = torch.from_numpy(np.ascontiguousarray(xyz.reshape(-1, 3))).cuda()
q = flat[:m].contiguous()
o = torch.tensor([flat.shape[0]], dtype=torch.int32, device="cuda")
qo = torch.tensor([m], dtype=torch.int32, device="cuda")
for _ in range(3):
    idx, d2 = _amloss.knn_raw(k, flat, q, o, qo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
idx, d2 = _amloss.knn_raw(k, flat, q, o, qo)
e1.record()
torch.cuda.synchronize()
print("knn", k, m, e0.elapsed_time(e1), "ms")
