import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
import test_gpu_quoted_configs as T
from oracle import ref_kernels as rk
from amcontrast3d_b200 import _amloss
from amcontrast3d_b200.replay import PathReplay
r = PathReplay(batch=2, n_points=64000, k=16, num_classes=20, ignore_index=-100, refine=True, refine_k=8,
               with_grouping=False, prefetch=False, loss_args=dict(temperature=0.5, nu=0.6))
loss = r.step(); torch.cuda.synchronize()
ref_loss, ref_a, inter, ref_g, p = T._oracle_on(r, r.f_dec, refine=True)
print("loss", loss.item(), ref_loss.item())
for s in range(4):
    g, rg = r.f_dec[s].grad, ref_g[s]
    err = (g - rg).norm(dim=1)
    print("stage", s, "rel", float((g - rg).norm() / rg.norm()), "rows with err>1e-3*max:", int((err > 1e-3 * rg.norm(dim=1).max()).sum()))
    w = torch.argsort(err, descending=True)[:6]
    pts = p[s].reshape(-1, 3).contiguous(); o = r._offsets[s]
    for K in (8, 16):
        i_ref, d_ref = rk.knnquery(K + 1, pts, pts, o, o)
        i_our, d_our = _amloss.knn_raw(K, pts, pts, o, o)
        neq = (i_ref[:, :K] != i_our).any(1)
        tie = ~(d_ref[:, 1:] > d_ref[:, :-1]).all(1)
        print("   K", K, "rows idx differ", int(neq.sum()), "tie rows", int(tie.sum()), "differ&~tie", int((neq & ~tie).sum()), "worst rows differ?", neq[w].tolist())
    print("   worst", w.tolist(), [float(x) for x in err[w]])
