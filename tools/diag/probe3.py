import sys, os, itertools
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes, _amloss
xyz, _ = scenes.batch_of_scenes(2, 4096, "volume", first_scene=3)
p = torch.from_numpy(xyz.reshape(-1, 3)).cuda()
o = torch.tensor([p.shape[0]], dtype=torch.int32, device="cuda")
f32 = lambda t: t.to(torch.float32)
for ke in (11, 15):
    idx, _ = _amloss.knn_raw(ke + 1, p, p, o, o)
    nidx = idx[:, 1:].long()
    for Bn in (1, 2, 3, 4, 8, 16, 32, 64, 100, 128, 256, 512, 4000):
        src = p[:Bn].unsqueeze(1); dst = p[nidx[:Bn]]
        mm = torch.matmul(src, dst.permute(0, 2, 1)).squeeze(1)
        a = [src[:, 0, c:c + 1].double() for c in range(3)]
        b = [dst[:, :, c].double() for c in range(3)]
        pr = [a[i] * b[i] for i in range(3)]
        res = {}
        for perm in itertools.permutations(range(3)):
            i, j, k = perm
            res[f"fma{perm}"] = f32(pr[k] + f32(pr[j] + f32(pr[i]).double()).double())
            res[f"sep{perm}"] = f32(f32(f32(pr[i]).double() + f32(pr[j]).double()).double() + f32(pr[k]).double())
            res[f"fma2sep{perm}"] = f32(f32(pr[j] + f32(pr[i]).double()).double() + f32(pr[k]).double())
            res[f"sepfma{perm}"] = f32(pr[k] + f32(f32(pr[i]).double() + f32(pr[j]).double()).double())
        best = sorted(((float((v == mm).float().mean()), k) for k, v in res.items()), reverse=True)[:2]
        sq = src ** 2; s3 = torch.sum(sq, -1)[:, 0]
        x, y, z = [sq[:, 0, i].double() for i in range(3)]
        c3 = {"(x+y)+z": f32(f32(x + y).double() + z), "x+(y+z)": f32(x + f32(y + z).double()), "(x+z)+y": f32(f32(x + z).double() + y)}
        b3 = sorted(((float((v == s3).float().mean()), k) for k, v in c3.items()), reverse=True)[:1]
        print(ke, Bn, best, b3)
