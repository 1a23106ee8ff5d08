import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes, _amloss
xyz, _ = scenes.batch_of_scenes(2, 4096, "volume", first_scene=3)
p = torch.from_numpy(xyz.reshape(-1, 3)).cuda()
o = torch.tensor([p.shape[0]], dtype=torch.int32, device="cuda")
f32 = lambda t: t.to(torch.float32)
for ke in (3, 7, 11, 15, 23, 31):
    idx, _ = _amloss.knn_raw(ke + 1, p, p, o, o)
    nidx = idx[:, 1:].long()
    kinds = []
    for Bn in range(120, 270):
        src = p[:Bn].unsqueeze(1); dst = p[nidx[:Bn]]
        mm = torch.matmul(src, dst.permute(0, 2, 1)).squeeze(1)
        a = [src[:, 0, c:c + 1].double() for c in range(3)]
        b = [dst[:, :, c].double() for c in range(3)]
        pr = [a[i] * b[i] for i in range(3)]
        big = f32(f32(pr[1] + f32(pr[0]).double()).double() + f32(pr[2]).double())
        small = f32(f32(f32(pr[2]).double() + f32(pr[0]).double()).double() + f32(pr[1]).double())
        kinds.append("B" if torch.equal(big, mm) else ("s" if torch.equal(small, mm) else "?"))
    s = "".join(kinds)
    print(ke, "first B at batch", 120 + s.index("B") if "B" in s else None, "pattern", s[:20], "...", s[-20:], "any ?", "?" in s)
