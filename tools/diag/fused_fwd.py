"""dev check of the fused forward against the torch composition (FP32, TF32 off) on the GPU"""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from amcontrast3d_b200 import scenes
from amcontrast3d_b200.layers import ball_query, furthest_point_sample, QueryAndGroup
from amcontrast3d_b200.layers.fused import FusedGroupConvBNReLUMax

def ref(q, p, f, idx, w, gamma, beta, radius, eps=1e-5):
    dp, fj = QueryAndGroup(radius, idx.shape[2], normalize_dp=True)(q, p, f, idx=idx)
    x = torch.cat([dp, fj], 1)
    y = torch.einsum("oc,bcps->bops", w, x)
    mean = y.mean(dim=(0, 2, 3), keepdim=True); var = y.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    z = (y - mean) / torch.sqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return torch.relu(z).max(-1)[0], mean.flatten(), var.flatten()

cases = [(2, 600, 600, 32, 32, 16, 0.2), (2, 400, 400, 64, 64, 32, 0.25), (2, 800, 200, 32, 64, 16, 0.15),
         (2, 2048, 2048, 128, 128, 32, 0.2), (8, 6000, 6000, 128, 128, 32, 0.2), (8, 24000, 6000, 64, 128, 32, 0.1),
         (8, 1500, 1500, 256, 256, 32, 0.4), (8, 375, 375, 512, 512, 32, 0.8), (8, 93, 93, 1024, 1024, 32, 1.6)]
only = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else None
for ci, (B, N, M, C, O, ns, radius) in enumerate(cases):
    if only is not None and ci not in only: continue
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
    idx = ball_query(radius, ns, p, q)
    r_out, r_mean, r_var = ref(q, p, f, idx, w, gamma, beta, radius)
    for prec in ("tf32x3", "tf32"):
        out, mean, var = FusedGroupConvBNReLUMax.apply(f, w, gamma, beta, q, p, idx, radius, True, 1e-5, prec)
        torch.cuda.synchronize()
        rel = lambda a, b: float((a - b).norm() / b.norm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): FusedGroupConvBNReLUMax.apply(f, w, gamma, beta, q, p, idx, radius, True, 1e-5, prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flop = 2.0 * B * M * ns * (C + 3) * O
        print(f"case {ci} B{B} N{N} M{M} C{C} O{O} ns{ns} {prec}: out rel {rel(out, r_out):.2e} mean rel {rel(mean, r_mean):.2e} var rel {rel(var, r_var):.2e}  {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s", flush=True)
