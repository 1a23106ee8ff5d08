"""A lean program for the per-kernel ncu captures of profiles/ (ncu saves and restores the memory a profiled kernel
writes on each of its ~40 passes, so the captured process must not hold the 6 GB of grouped tensors a whole step
produces):  the config-2 geometry + loss path WITHOUT the grouping tensors (FPS chain, self / label kNN, three_nn,
loss forward), one ball query per level, and ONE grouping gather + scatter-add and one three_interpolate pair at
their largest config-2 shape.

    python tools/ncu_targets.py > plain.log && ncu --set full --clock-control none \
        -k regex:'knn_wq_kernel|knn_tq_kernel|ball_wq_kernel|amloss_forward_kernel|fps_cluster_kernel|group_fwd_tma_kernel|group_bwd_tma_kernel|interp_fwd_tma|interp_bwd_tma' \
        -s <launches of pass 1> -o gpurun_out/r02_kernels python tools/ncu_targets.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from amcontrast3d_b200.layers import ball_query, grouping_operation, three_interpolation  # noqa: E402
from amcontrast3d_b200.replay import PathReplay  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
r = PathReplay(batch=8, n_points=24000, k=16, with_grouping=False, geometry_stream=False, prefetch=False)
p = r._fps_chain(r.d_xyz)
f = torch.randn(8, 128, 6000, device="cuda", requires_grad=True)
fc = torch.randn(8, 128, 6000, device="cuda", requires_grad=True)
for _ in range(passes):
    loss = r.step()                                            # 4 FPS, 7 kNN, 4 three_nn + interpolate, loss
    for l in range(1, 5):                                      # the two ball-query shapes of every level
        ball_query(0.1 * 2 ** (l - 1), 32, p[l - 1], p[l])
        idx = ball_query(0.1 * 2 ** l, 32, p[l], p[l])
        if l == 1:
            idx1 = idx
    out = grouping_operation(f, idx1)                          # (8,128,6000) x (6000,32): the largest feature grouping
    out.backward(torch.ones_like(out))
    up = three_interpolation(p[0], p[1], fc)                   # 24000 <- 6000, C = 128
    up.backward(torch.ones_like(up))
    del out, up
torch.cuda.synchronize()
print("loss", float(loss))
