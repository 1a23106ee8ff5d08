import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch
from oracle import loss_oracle as lo, ref_kernels as rk
from amcontrast3d_b200 import _amloss
from amcontrast3d_b200.replay import PathReplay
from amcontrast3d_b200.AMContrast3D.MarginContrast import _stage_ambiguity
from amcontrast3d_b200.AMContrast3D.AEF.utils import get_ftype
r = PathReplay(batch=2, n_points=64000, k=16, num_classes=20, ignore_index=-100, refine=False,
               with_grouping=False, prefetch=False, loss_args=dict(temperature=0.5, nu=0.6))
p = r._fps_chain(r.d_xyz)
sl = lo.make_stage_list([p[s].reshape(-1, 3).contiguous() for s in range(4)], [f.detach() for f in r.f_dec])
nstride = torch.tensor([4, 4, 4, 4])
ref = lo.stage_prelude("up", 0, sl, r.d_labels.reshape(-1), 20, -100, r.args, nstride, rk.knnquery)
st = _stage_ambiguity("up", 0, sl, r.d_labels.reshape(-1), 20, -100, r.args, nstride, get_ftype("latent")[0])
a, ra = st["a"], ref["a"]
d = (a - ra).abs()
print("max |da|", float(d.max()), "n != ", int((a != ra).sum()), "n rel>2e-6", int((d > 2e-6 * ra.abs()).sum()))
bad = torch.nonzero(d > 2e-6 * ra.abs()).flatten()[:10]
P = p[0].reshape(-1, 3)
pm = ref["posmask"]; nidx = ref["neighbor_idx"].long()
for i in bad.tolist():
    BNC = P[i].view(1, 1, 3); BMC = P[nidx[i]].view(1, -1, 3)
    dd = lo.square_distance(BNC, BMC).squeeze()
    m = pm[i].int()
    dpos = torch.sum(m * dd, -1); dneg = torch.sum((1 - m) * dd, -1)
    # sequential in f64 -> f32
    seqp = torch.zeros((), device="cuda"); seqn = torch.zeros((), device="cuda")
    for j in range(15):
        if m[j]: seqp = (seqp.double() + dd[j].double()).float()
        else: seqn = (seqn.double() + dd[j].double()).float()
    print(i, "ours", float(a[i]), "ref", float(ra[i]), "npos", int(m.sum()), "dpos %.9g seq %.9g dneg %.9g seq %.9g" % (float(dpos), float(seqp), float(dneg), float(seqn)), "dd min %.3g" % float(dd.min()), "|p|", float(P[i].norm()))
print("---- per-neighbour, row 86558")
i = 86558
f32 = lambda t: t.to(torch.float32)
src = P[i].view(1, 1, 3); dst = P[nidx[i]].view(1, -1, 3)
x1, y1, z1 = [src[:, 0, c:c + 1].double() for c in range(3)]
x2, y2, z2 = [dst[:, :, c].double() for c in range(3)]
dot = f32(f32(y1 * y2 + f32(x1 * x2).double()).double() + f32(z1 * z2).double())
n1 = f32(f32(f32(x1 * x1).double() + f32(z1 * z1).double()).double() + f32(y1 * y1).double())
n2 = f32(f32(f32(x2 * x2).double() + f32(z2 * z2).double()).double() + f32(y2 * y2).double())
em = f32(f32(f32(-2.0 * dot.double()).double() + n1.double()).double() + n2.double())
# torch on the full boundary batch, as the oracle runs it
a0 = ref["a"]; pmk = ref["posmask"]
mask_num = pmk.int().sum(-1); mxx = mask_num.max()
idx_b = (mask_num > 0) & (mask_num < mxx)
rows = torch.nonzero(idx_b).flatten()
pos = int((rows == i).nonzero())
BNC = P[idx_b].unsqueeze(1); BMC = P[nidx[idx_b]]
dd_full = lo.square_distance(BNC, BMC)[pos, 0]
dd_one = lo.square_distance(src, dst)[0, 0]
print("torch full batch:", dd_full.tolist())
print("torch one row   :", dd_one.tolist())
print("emulated        :", em[0].tolist())
print("posmask         :", pmk[i].int().tolist())
print("true f64        :", ((P[i].double() - P[nidx[i]].double()) ** 2).sum(-1).tolist())
print("n boundary", int(idx_b.sum()), "pos", pos)
mm_full = torch.matmul(BNC, BMC.permute(0, 2, 1))[pos, 0]
print("mm full", mm_full.tolist()); print("dot emu", dot[0].tolist())
s1 = torch.sum(BNC ** 2, -1)[pos]; print("n1 full", s1.tolist(), "emu", n1.tolist())
s2 = torch.sum(BMC ** 2, -1)[pos]; print("n2 full", s2.tolist()); print("n2 emu ", n2[0].tolist())
