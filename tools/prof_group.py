"""grouping_operation forward / backward at the shapes of config 2 with real ball-query indices:
checks against torch.gather / index_add and prints achieved GB/s (algorithmic bytes / CUDA-event time).

    AMC3D_GROUP_IMPL=0|1|2 python tools/prof_group.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import ball_query, furthest_point_sample, grouping_operation

B = 8
xyz, _ = scenes.batch_of_scenes(B, 24000, "surface")
p0 = torch.from_numpy(xyz).cuda()
levels = [p0]
for n in (6000, 1500, 375, 93):
    i = furthest_point_sample(levels[-1], n).long()
    levels.append(torch.gather(levels[-1], 1, i.unsqueeze(-1).expand(-1, -1, 3)).contiguous())

# (support level, query level, C, radius)
shapes = [(0, 1, 64, 0.1), (1, 1, 128, 0.2), (1, 2, 128, 0.2), (2, 2, 256, 0.4), (3, 3, 512, 0.8), (4, 4, 1024, 1.6)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else -1
tot_f = tot_b = 0.0
for si, (ls, lq, C, r) in enumerate(shapes):
    if only >= 0 and si != only:
        continue
    sup, qry = levels[ls], levels[lq]
    N, M = sup.shape[1], qry.shape[1]
    idx = ball_query(r, 32, sup, qry)
    f = torch.randn(B, C, N, device="cuda", requires_grad=True)
    out = grouping_operation(f, idx)
    ref = torch.gather(f.detach(), 2, idx.reshape(B, 1, -1).expand(-1, C, -1).long()).reshape(out.shape)
    assert torch.equal(out.detach(), ref), "forward mismatch"
    g = torch.randn_like(out)
    out.backward(g)
    gref = torch.zeros(B, C, N, device="cuda").scatter_add_(2, idx.reshape(B, 1, -1).expand(-1, C, -1).long(),
                                                           g.reshape(B, C, -1))
    err = ((f.grad - gref).norm() / gref.norm()).item()
    assert err < 1e-5, err
    del ref, gref
    reps = 5
    tf = tb = 0.0
    for it in range(reps + 2):
        f.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        out = grouping_operation(f, idx)
        e[1].record()
        e[2].record()
        out.backward(g)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf += e[0].elapsed_time(e[1]) / reps
            tb += e[2].elapsed_time(e[3]) / reps
    P = M * 32
    bytes_f = 4.0 * B * (C * P + C * N + P)
    bytes_b = 4.0 * B * (C * P + 2 * C * N + P)
    tot_f += tf
    tot_b += tb
    print(f"C={C:5d} N={N:6d} M={M:5d}: fwd {tf:.4f} ms {bytes_f / tf / 1e6:7.1f} GB/s | bwd {tb:.4f} ms "
          f"{bytes_b / tb / 1e6:7.1f} GB/s | grad rel err {err:.2e}")
print(f"impl={os.environ.get('AMC3D_GROUP_IMPL', 'default')} sum fwd {tot_f:.4f} ms bwd {tot_b:.4f} ms")
