"""Two serial (single-stream, eager) path-replay steps at BASELINE config 2 — the program the per-kernel ncu
captures of profiles/ are taken on:

    python tools/one_step.py > plain.log && ncu --set full --clock-control none --import-source on \
        -k regex:'knn_wq_kernel|knn_tq_kernel|ball_wq_kernel|amloss_forward_kernel|fps_cluster_kernel|group_fwd_tma_kernel|group_bwd_tma_kernel|interp_' \
        -s <launches of step 1> -c <launches of step 2> -o gpurun_out/r02_step python tools/one_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from amcontrast3d_b200.replay import PathReplay  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
r = PathReplay(batch=8, n_points=24000, k=16, geometry_stream=False, prefetch=False)
for _ in range(steps):
    loss = r.step()
torch.cuda.synchronize()
print("loss", float(loss))
