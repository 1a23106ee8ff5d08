import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import numpy as np, torch
from oracle import loss_oracle as lo, ref_kernels as rk
from amcontrast3d_b200.replay import PathReplay
from amcontrast3d_b200 import _amloss
from amcontrast3d_b200.AMContrast3D.MarginContrast import _stage_ambiguity
from amcontrast3d_b200.AMContrast3D.AEF.utils import get_ftype

r = PathReplay(batch=8, n_points=24000, k=16, with_grouping=False, prefetch=False, geometry_stream=False)
p = r._fps_chain(r.d_xyz)
sl = lo.make_stage_list([p[s].reshape(-1, 3).contiguous() for s in range(4)], [f.detach().clone().double().requires_grad_(True) for f in r.f_dec])
l64, _, _, inter = lo.contrast_head_forward(r.d_labels.reshape(-1), sl, 13, None, r.args, knn=rk.knnquery)
l64.backward()
g64 = [d["f_out"].grad for d in sl["up"]]
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
nstride = torch.tensor([4, 4, 4, 4])
slg = {"down": [{"p_out": p[s].reshape(-1, 3).contiguous(), "offset": r._offsets[s]} for s in range(4)]}
slg["up"] = slg["down"]
for s in (0, 1):
    st = _stage_ambiguity("up", s, slg, r.d_labels.reshape(-1), 13, None, r.args, nstride, get_ftype("latent")[0])
    ref = inter[s]
    print("stage", s, "idx equal rows", float((st["knn_idx"] == ref["knn_idx"]).all(1).float().mean()), "a maxdiff", float((st["a"] - ref["a"]).abs().max()),
          "sel equal", bool(torch.equal((st["a"] > 0) & (st["a"] <= 1), ref["sel"])), "stats", st["stats"].tolist())
    for name, order in (("none", None), ("order", st["order"]), ("compact", _amloss.compact_order(st["order"], st["a"]))):
        f = r.f_dec[s].detach().clone().requires_grad_(True)
        loss = _amloss.am_loss(f, st["nl"], st["posbits"], st["a"], st["stats"], r.args, order)
        loss.backward()
        torch.cuda.synchronize()
        print("   order=%s loss %.9f  grad rel err vs f64 %.3e" % (name, loss.item(), rel(f.grad, g64[s])))
    # per-row loss
    f = r.f_dec[s].detach().clone().requires_grad_(True)
    m, d = f.shape
    import ctypes
    from amcontrast3d_b200 import _capi
    from amcontrast3d_b200._capi import ptr, stream
    inv = torch.empty(m, device="cuda"); loss_pt = torch.zeros(m, device="cuda"); ghat = torch.zeros(m, d, device="cuda")
    prm = _amloss.loss_params(r.args)
    _capi.call("amc3d_row_inv_norm", m, d, ptr(f), ptr(inv), stream(f))
    _capi.call("amc3d_amloss_forward_order", m, d, st["nl"].ke, st["nl"].ld, ptr(f), ptr(inv), st["nl"].ptr, ptr(st["posbits"]), ptr(st["a"]), ctypes.byref(prm), ptr(loss_pt), ptr(ghat), 0, stream(f))
    torch.cuda.synchronize()
    lr = ref["loss_rows"].detach()
    mine = loss_pt[ref["sel"]]
    print("   loss rows: rel err %.3e max abs %.3e" % (rel(mine, lr), float((mine.double() - lr).abs().max())))
    # ghat reference in f64: d(sum of loss rows)/du
    fd = r.f_dec[s].detach().double()
    u = (fd / fd.norm(dim=1, keepdim=True).clamp_min(1e-8)).requires_grad_(True)
    nidx = ref["neighbor_idx"].long()
    sel = ref["sel"]
    un = u[nidx[sel].reshape(-1)].view(int(sel.sum()), nidx.shape[1], d)
    dist = (u[sel].unsqueeze(1) * un).sum(-1)
    rows = lo.contrast_softnn_margin(dist, ref["posmask"][sel], ref["a"][sel].double(), r.args)
    rows.sum().backward()
    print("   ghat rel err vs f64 %.3e ; ghat max abs err %.3e" % (rel(ghat, u.grad), float((ghat.double() - u.grad).abs().max())))
    rowerr = (ghat.double() - u.grad).norm(dim=1)
    w = torch.argsort(rowerr, descending=True)[:8]
    for i in w.tolist():
        print("      row", i, "err %.3e |ghat| %.3e sel %d  n_as_neighbour %d" % (float(rowerr[i]), float(u.grad[i].norm()), int(sel[i]), int((nidx[sel] == i).sum())))
