import sys, os, itertools
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes, _amloss
xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
p = torch.from_numpy(xyz.reshape(-1, 3)).cuda()
o = torch.tensor([p.shape[0]], dtype=torch.int32, device="cuda")
f32 = lambda t: t.to(torch.float32)
for ke, Bn in ((15, 60000), (15, 82307), (15, 1000), (23, 60000)):
    idx, _ = _amloss.knn_raw(ke + 1, p, p, o, o)
    nidx = idx[:, 1:].long()
    src = p[:Bn].unsqueeze(1); dst = p[nidx[:Bn]]
    mm = torch.matmul(src, dst.permute(0, 2, 1)).squeeze(1)
    a = [src[:, 0, c:c + 1].double() for c in range(3)]
    b = [dst[:, :, c].double() for c in range(3)]
    pr = [a[i] * b[i] for i in range(3)]          # exact products in f64
    res = {}
    for perm in itertools.permutations(range(3)):
        i, j, k = perm
        res[f"fma{perm}"] = f32(pr[k] + f32(pr[j] + f32(pr[i]).double()).double())
        res[f"sep{perm}"] = f32(f32(f32(pr[i]).double() + f32(pr[j]).double()).double() + f32(pr[k]).double())
        res[f"fma2sep{perm}"] = f32(f32(pr[j] + f32(pr[i]).double()).double() + f32(pr[k]).double())   # fma(i,j) + rounded k
        res[f"sepfma{perm}"] = f32(pr[k] + f32(f32(pr[i]).double() + f32(pr[j]).double()).double())
    best = sorted(((float((v == mm).float().mean()), k) for k, v in res.items()), reverse=True)[:5]
    print(ke, Bn, best)
    # is bmm deterministic wrt batch size / position?
    mm2 = torch.matmul(src[:500], dst[:500].permute(0, 2, 1)).squeeze(1)
    print("   same as small batch:", bool(torch.equal(mm2, mm[:500])))
    # einsum / elementwise alternative the reference does not use, for reference
    el = (src * dst).sum(-1)
    print("   (src*dst).sum(-1) == matmul:", float((el == mm).float().mean()))
    # mismatch pattern vs column index
    best_name = best[0][1]
    mis = (res[best_name] != mm).float().mean(0)
    print("   mismatch by column:", [round(float(x), 3) for x in mis])
