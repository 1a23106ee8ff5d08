"""GPU-kernel breakdown (CUPTI via torch.profiler) of the fused operator's forward + backward over config 2's 19 layers"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch, torch.nn as nn
from torch.profiler import profile, ProfilerActivity
import bench
from amcontrast3d_b200.replay import PathReplay
from amcontrast3d_b200.layers import ball_query
from amcontrast3d_b200.layers.fused import fused_group_conv_bn_relu_max
dev = torch.device("cuda", 0)
replay = PathReplay(batch=8, n_points=24000, device=dev, k=16)
p = replay._fps_chain(replay.d_xyz)
g = torch.Generator(device=dev); g.manual_seed(5)
work = []
for kind, l, N, M, cin, cout, r in bench.xl_layers(replay.B, replay.N):
    sup, qry = (p[l - 1], p[l]) if kind == "sa" else (p[l], p[l])
    idx = ball_query(r, 32, sup, qry)
    conv = nn.Conv2d(cin + 3, cout, 1, bias=False).to(dev); bn = nn.BatchNorm2d(cout).to(dev)
    f = torch.randn((replay.B, cin, N), device=dev, generator=g).requires_grad_(True)
    go = torch.randn((replay.B, cout, M), device=dev, generator=g)
    work.append((qry, sup, idx, conv, bn, f, go, r))
def run(backward=True):
    for qry, sup, idx, conv, bn, f, go, r in work:
        out = fused_group_conv_bn_relu_max(qry, sup, f, idx, conv.weight, bn, r, True, "tf32")
        if backward: out.backward(go)
for _ in range(2): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda x: -x[1])
tot = sum(r[1] for r in rows)
print(f"total GPU kernel time {tot / 1e3:.3f} ms over {sum(r[2] for r in rows)} launches")
for k, t, c in rows[:40]:
    print(f"{t / 1e3:8.3f} ms {c:5d}  {k[:140]}")
