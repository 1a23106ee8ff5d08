"""where the host time of the fused operator's eager forward + backward goes (cProfile over the 19 config-2 layers)"""
import os, sys, cProfile, pstats, io
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch, torch.nn as nn
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
def run():
    for qry, sup, idx, conv, bn, f, go, r in work:
        out = fused_group_conv_bn_relu_max(qry, sup, f, idx, conv.weight, bn, r, True, "tf32")
        out.backward(go)
for _ in range(3): run()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5): run()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per pass {1e3 * (t1 - t0) / 5:.2f} ms (enqueue only), {1e3 * (t2 - t0) / 5:.2f} ms incl. drain")
pr = cProfile.Profile(); pr.enable()
for _ in range(5): run()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
