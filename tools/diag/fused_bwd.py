"""dev check of the fused backward against autograd over the torch composition in FP64 on the GPU"""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import ball_query, furthest_point_sample, QueryAndGroup
from amcontrast3d_b200.layers.fused import FusedGroupConvBNReLUMax

def ref(q, p, f, idx, w, gamma, beta, radius, eps=1e-5):
    dp, fj = QueryAndGroup(radius, idx.shape[2], normalize_dp=True)(q, p, f.float(), idx=idx)
    B, C, N = f.shape
    fj = torch.gather(f, 2, idx.reshape(B, 1, -1).expand(-1, C, -1).long()).reshape(B, C, idx.shape[1], idx.shape[2])
    x = torch.cat([dp.to(f.dtype), fj], 1)
    y = torch.einsum("oc,bcps->bops", w, x)
    mean = y.mean(dim=(0, 2, 3), keepdim=True); var = y.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    z = (y - mean) / torch.sqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return torch.relu(z).max(-1)[0]

cases = [(2, 600, 600, 32, 32, 16, 0.2), (2, 400, 400, 64, 64, 32, 0.25), (2, 800, 200, 32, 64, 16, 0.15),
         (2, 2048, 2048, 128, 128, 32, 0.2), (8, 6000, 6000, 128, 128, 32, 0.2), (8, 24000, 6000, 64, 128, 32, 0.1),
         (8, 1500, 1500, 256, 256, 32, 0.4), (8, 375, 375, 512, 512, 32, 0.8), (8, 93, 93, 1024, 1024, 32, 1.6)]
sel = [int(x) for x in sys.argv[1:]] or list(range(len(cases)))
for ci in sel:
    B, N, M, C, O, ns, radius = cases[ci]
    xyz, _ = scenes.batch_of_scenes(B, N, "surface", first_scene=31)
    p = torch.from_numpy(xyz).cuda()
    if M == N: q = p
    else:
        i = furthest_point_sample(p, M).long(); q = torch.gather(p, 1, i.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    g = torch.Generator(device="cuda").manual_seed(17)
    f = torch.randn(B, C, N, device="cuda", generator=g)
    w = torch.randn(O, C + 3, device="cuda", generator=g) / (C + 3) ** 0.5
    gamma = 1 + 0.1 * torch.randn(O, device="cuda", generator=g); gamma[::7] *= -1
    beta = 0.1 * torch.randn(O, device="cuda", generator=g)
    go = torch.randn(B, O, M, device="cuda", generator=g)
    idx = ball_query(radius, ns, p, q)
    leaves = [t.double().requires_grad_(True) for t in (f, w, gamma, beta)]
    r_out = ref(q, p, *leaves[:1], idx, *leaves[1:], radius)
    r_out.backward(go.double())
    rg = [t.grad for t in leaves]
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    for prec in ("tf32x3", "tf32"):
        mine = [t.clone().requires_grad_(True) for t in (f, w, gamma, beta)]
        out, mean, var = FusedGroupConvBNReLUMax.apply(mine[0], mine[1], mine[2], mine[3], q, p, idx, radius, True, 1e-5, prec)
        out.backward(go)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outs = [FusedGroupConvBNReLUMax.apply(mine[0], mine[1], mine[2], mine[3], q, p, idx, radius, True, 1e-5, prec)[0] for _ in range(3)]
        torch.cuda.synchronize()
        e0.record()
        for o_ in outs: o_.backward(go)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"case {ci} C{C} O{O} ns{ns} M{M} {prec}: out {rel(out, r_out):.1e} df {rel(mine[0].grad / 4, rg[0]):.1e} dW {rel(mine[1].grad / 4, rg[1]):.1e} "
              f"dgamma {rel(mine[2].grad / 4, rg[2]):.1e} dbeta {rel(mine[3].grad / 4, rg[3]):.1e}  bwd {ms:.3f} ms", flush=True)
