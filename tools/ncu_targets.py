"""ONE launch of each kernel the per-kernel ncu summaries of profiles/ are taken on, at its BASELINE config-2 shape
(ncu replays a profiled kernel ~40 times and saves / restores the memory it writes on every pass, so the captured
process holds nothing it does not need):

    python tools/ncu_targets.py > plain.log && ncu --set full --clock-control none \
        -k regex:'knn_wq_kernel|knn_tq_kernel|ball_wq_kernel|amloss_forward_kernel|fps_cluster_kernel|group_fwd_tma_kernel|group_bwd_tma_kernel|fused_sa_fwd_kernel' \
        -o /tmp/r02_kernels python tools/ncu_targets.py
    ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv      # small; summarised by
    python tools/ncu_kernel_table.py gpurun_out/r02_kernels_raw.csv > profiles/r02_kernels_ncu.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from amcontrast3d_b200 import _amloss, scenes  # noqa: E402
from amcontrast3d_b200.layers import ball_query, furthest_point_sample, grouping_operation, three_nn  # noqa: E402
from amcontrast3d_b200.layers.fused import FusedGroupConvBNReLUMax  # noqa: E402
from amcontrast3d_b200.replay import aa_args  # noqa: E402

xyz, lab = scenes.batch_of_scenes(8, 24000, "surface")
p0 = torch.from_numpy(xyz).cuda()
labels = torch.from_numpy(lab).cuda().reshape(-1)
idx = furthest_point_sample(p0, 6000)                                                  # fps_cluster_kernel<16,12,4>
p1 = torch.gather(p0, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
flat = p0.reshape(-1, 3).contiguous()
o = torch.tensor([flat.shape[0]], dtype=torch.int32, device="cuda")
knn_idx, _, order = _amloss.knn_raw(16, flat, flat, o, o, want_order=True)              # knn_wq_kernel<1>: 192 000 self queries
three_nn(p0, p1)                                                                       # knn_tq_kernel<3>
bq = ball_query(0.2, 32, p1, p1)                                                       # ball_wq_kernel<1>
cls, _ = _amloss.stage_labels(labels, 13, None)
nl = _amloss.NeighbourList(knn_idx, drop_self=True)
posbits, cnt, mx = _amloss.posmask_count(nl, cls)
a, stats = _amloss.ambiguity(flat, nl, posbits, cnt, mx, "Method2", 0.04, 0.5)
f0 = torch.randn(flat.shape[0], 64, device="cuda", requires_grad=True)
loss = _amloss.am_loss(f0, nl, posbits, a, stats, aa_args(16), _amloss.compact_order(order, a))   # amloss_forward_kernel<16,1>
f = torch.randn(8, 128, 6000, device="cuda", requires_grad=True)
out = grouping_operation(f, bq)                                                        # group_fwd_tma_kernel<4>
out.backward(torch.ones_like(out))                                                     # group_bwd_tma_kernel<1>
del out
w = torch.randn(128, 131, device="cuda") / 11.4
y, _, _ = FusedGroupConvBNReLUMax.apply(f.detach(), w, torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"),
                                        p1, p1, bq, 0.2, True, 1e-5, "tf32")            # fused_sa_fwd_kernel<32,false>
torch.cuda.synchronize()
print("loss", float(loss), "fused mean", float(y.mean()))
