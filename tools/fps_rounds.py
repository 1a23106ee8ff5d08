"""AMC3D_FPS_DEBUG=1 python tools/fps_rounds.py : exchange rounds per FPS call along the PointNeXt chain
(the kernel then leaves the round count in temp[0]); picks per round = (m - 1) / rounds."""
import os
import sys

os.environ["AMC3D_FPS_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200 import pointnet2_batch_cuda as ext
from amcontrast3d_b200 import scenes

xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
p = torch.from_numpy(xyz).cuda()
while p.shape[1] >= 4 * 93:
    B, n, _ = p.shape
    m = n // 4
    temp = torch.full((B, n), 1e10, device="cuda")
    idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    ext.furthest_point_sampling_wrapper(B, n, m, p, temp, idx)
    r = temp[:, 0].cpu().tolist()
    inorder = bool((idx[0].cpu() == torch.arange(m, dtype=torch.int32)).all())
    print(f"n={n} m={m}: rounds per scene {r[:4]} -> {(m - 1) / (sum(r) / len(r)):.2f} picks/round; picks in index order: {inorder}")
    p = torch.gather(p, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
