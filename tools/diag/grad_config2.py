"""Diagnostic: where do the config-2 loss gradients differ from the torch oracle (fp32 and fp64)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from oracle import loss_oracle as lo, ref_kernels as rk
from amcontrast3d_b200.replay import PathReplay

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
r = PathReplay(batch=8, n_points=24000, k=k, with_grouping=False, prefetch=False)
loss = r.step(); torch.cuda.synchronize()
p = r._fps_chain(r.d_xyz)
def run(dtype):
    f_in = [f.detach().clone().to(dtype).requires_grad_(True) for f in r.f_dec]
    sl = lo.make_stage_list([p[s].reshape(-1, 3).contiguous() for s in range(4)], f_in)
    l, a_cat, a_list, inter = lo.contrast_head_forward(r.d_labels.reshape(-1), sl, 13, None, r.args, knn=rk.knnquery)
    l.backward()
    return l, inter, [f.grad for f in f_in]
l32, inter, g32 = run(torch.float32)
l64, _, g64 = run(torch.float64)
print("loss ours %.9f ref32 %.9f ref64 %.12f" % (loss.item(), l32.item(), l64.item()))
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for s in range(4):
    go = r.f_dec[s].grad
    print(f"stage {s}: ours-vs-32 {rel(go, g32[s]):.3e}  ours-vs-64 {rel(go, g64[s]):.3e}  32-vs-64 {rel(g32[s], g64[s]):.3e}  nsel {int(inter[s]['sel'].sum())}")
    rn = (go.double() - g64[s]).norm(dim=1); bn = g64[s].norm(dim=1)
    rr = rn / bn.clamp_min(1e-30)
    live = bn > 0
    print("   rows live", int(live.sum()), "row rel err ours-vs-64: median %.2e max %.2e  #>1e-4: %d" % (float(rr[live].median()), float(rr[live].max()), int((rr[live] > 1e-4).sum())))
    rn2 = (g32[s].double() - g64[s]).norm(dim=1); rr2 = rn2 / bn.clamp_min(1e-30)
    print("   row rel err ref32-vs-64: median %.2e max %.2e  #>1e-4: %d" % (float(rr2[live].median()), float(rr2[live].max()), int((rr2[live] > 1e-4).sum())))
    worst = torch.argsort(rn, descending=True)[:5]
    for w in worst.tolist():
        print("   worst row", w, "abs err ours %.3e ref32 %.3e norm %.3e sel %d a %.6f" % (float(rn[w]), float(rn2[w]), float(bn[w]), int(inter[s]['sel'][w]), float(inter[s]['a'][w])))
    # our stats
from amcontrast3d_b200.AMContrast3D.MarginContrast import _stage_ambiguity
