"""per-entry-point time of the fused operator's forward + backward at one PointNeXt-XL layer shape"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes, _capi
from amcontrast3d_b200.layers import ball_query, furthest_point_sample
from amcontrast3d_b200.layers.fused import FusedGroupConvBNReLUMax
cases = {"l1sa": (8, 24000, 6000, 64, 128, 32, 0.1), "l1la": (8, 6000, 6000, 128, 128, 32, 0.2), "l2la": (8, 1500, 1500, 256, 256, 32, 0.4),
         "l3la": (8, 375, 375, 512, 512, 32, 0.8), "l4la": (8, 93, 93, 1024, 1024, 32, 1.6)}
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
for name in (sys.argv[2:] or list(cases)):
    B, N, M, C, O, ns, radius = cases[name]
    xyz, _ = scenes.batch_of_scenes(B, N, "surface", first_scene=31)
    p = torch.from_numpy(xyz).cuda()
    q = p
    if M != N:
        i = furthest_point_sample(p, M).long(); q = torch.gather(p, 1, i.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    g = torch.Generator(device="cuda").manual_seed(17)
    f = torch.randn(B, C, N, device="cuda", generator=g).requires_grad_(True)
    w = (torch.randn(O, C + 3, device="cuda", generator=g) / (C + 3) ** 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(O, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(O, device="cuda", generator=g)).requires_grad_(True)
    go = torch.randn(B, O, M, device="cuda", generator=g)
    idx = ball_query(radius, ns, p, q)
    def run():
        out, _, _ = FusedGroupConvBNReLUMax.apply(f, w, gamma, beta, q, p, idx, radius, True, 1e-5, prec)
        return out
    for _ in range(2): run().backward(go)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    _capi.PROFILE = []
    ev[0].record(); out = run(); ev[1].record(); out.backward(go); ev[2].record()
    torch.cuda.synchronize()
    prof, _capi.PROFILE = _capi.PROFILE, None
    flop = 2.0 * B * M * ns * (C + 3) * O
    print(f"{name} {prec}: fwd {ev[0].elapsed_time(ev[1]):.3f} ms  bwd {ev[1].elapsed_time(ev[2]):.3f} ms   conv GFLOP {flop / 1e9:.1f}")
    for n_, e0, e1, a in prof:
        print(f"     {n_:36s} {e0.elapsed_time(e1):.3f} ms")
