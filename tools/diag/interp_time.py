"""per-call time of three_interpolate forward / backward at the four decoder shapes of config 2"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes, _capi
from amcontrast3d_b200.layers import three_nn, three_interpolate, furthest_point_sample
xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
p = [torch.from_numpy(xyz).cuda()]
for m in (6000, 1500, 375, 93):
    i = furthest_point_sample(p[-1], m).long()
    p.append(torch.gather(p[-1], 1, i.unsqueeze(-1).expand(-1, -1, 3)).contiguous())
tot = {"amc3d_three_interpolate_ws": 0.0, "amc3d_three_interpolate_grad_ws_set": 0.0}
for l, C in ((0, 128), (1, 256), (2, 512), (3, 1024)):
    d, idx = three_nn(p[l], p[l + 1])
    w = 1.0 / (d + 1e-8); w = (w / w.sum(-1, keepdim=True)).contiguous()
    f = torch.randn(8, C, p[l + 1].shape[1], device="cuda", requires_grad=True)
    go = torch.randn(8, C, p[l].shape[1], device="cuda")
    for _ in range(3): three_interpolate(f, idx, w).backward(go)
    torch.cuda.synchronize()
    _capi.PROFILE = []
    for _ in range(5): three_interpolate(f, idx, w).backward(go)
    torch.cuda.synchronize()
    prof, _capi.PROFILE = _capi.PROFILE, None
    by = {}
    for n_, e0, e1, a in prof: by.setdefault(n_, []).append(e0.elapsed_time(e1))
    line = f"level {l} C={C} n={p[l].shape[1]} m={p[l+1].shape[1]}: "
    for n_, v in by.items():
        line += f"{n_} {min(v)*1e3:.1f} us  "
        if n_ in tot: tot[n_] += min(v)
    print(line)
print({k: round(v, 4) for k, v in tot.items()})
