"""Which FP32 evaluation order do torch's CUDA kernels use for the pieces of AEF/function.py square_distance
and the 15-wide masked sums of AEF/ambiguity.py?  Candidates are emulated through float64."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R)
import torch
from amcontrast3d_b200 import scenes, _amloss
xyz, _ = scenes.batch_of_scenes(8, 24000, "surface")
p = torch.from_numpy(xyz.reshape(-1, 3)).cuda()
o = torch.tensor([p.shape[0]], dtype=torch.int32, device="cuda")
for ke in (15, 23):
    idx, _ = _amloss.knn_raw(ke + 1, p, p, o, o)
    nidx = idx[:, 1:].long()
    Bn = 60000
    src = p[:Bn].unsqueeze(1)                      # (B,1,3)
    dst = p[nidx[:Bn]]                             # (B,ke,3)
    mm = torch.matmul(src, dst.permute(0, 2, 1)).squeeze(1)     # (B,ke)
    f32 = lambda t: t.to(torch.float32)
    x1, y1, z1 = [src[:, 0, c:c + 1].double() for c in range(3)]
    x2, y2, z2 = [dst[:, :, c].double() for c in range(3)]
    cands = {
        "sep (x+y)+z": f32(f32(f32(x1 * x2).double() + f32(y1 * y2).double()).double() + f32(z1 * z2).double()),
        "fma asc": f32(z1 * z2 + f32(y1 * y2 + f32(x1 * x2).double()).double()),
        "fma desc": f32(x1 * x2 + f32(y1 * y2 + f32(z1 * z2).double()).double()),
        "exact": f32(x1 * x2 + y1 * y2 + z1 * z2),
    }
    print(f"ke={ke} matmul:", {k: round(float((v == mm).float().mean()), 5) for k, v in cands.items()})
    sq = src ** 2
    s3 = torch.sum(sq, -1)                         # (B,1)
    a, b, c = [sq[:, 0, i].double() for i in range(3)]
    c3 = {"(a+b)+c": f32(f32(a + b).double() + c), "a+(b+c)": f32(a + f32(b + c).double()), "(a+c)+b": f32(f32(a + c).double() + b)}
    print("   sum3 src:", {k: round(float((v == s3[:, 0]).float().mean()), 5) for k, v in c3.items()})
    sq2 = dst ** 2
    s3d = torch.sum(sq2, -1)
    a, b, c = [sq2[:, :, i].double() for i in range(3)]
    c3 = {"(a+b)+c": f32(f32(a + b).double() + c), "a+(b+c)": f32(a + f32(b + c).double()), "(a+c)+b": f32(f32(a + c).double() + b)}
    print("   sum3 dst:", {k: round(float((v == s3d).float().mean()), 5) for k, v in c3.items()})
    # masked sum over ke
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1)); dist += s3.view(Bn, 1, 1); dist += s3d.view(Bn, 1, ke)
    mask = (torch.rand(Bn, ke, device="cuda") < 0.5).int()
    v = mask * dist.squeeze()
    tot = torch.sum(v, -1)
    vd = v.double()
    seq = torch.zeros(Bn, device="cuda")
    for j in range(ke):
        seq = f32(seq.double() + vd[:, j])
    def tree(vals, n):
        w = 1
        while w < n: w *= 2
        buf = torch.zeros(Bn, w, device="cuda", dtype=torch.float64); buf[:, :n] = vals
        off = 1
        buf = buf.clone()
        while off < w:
            nb = buf.clone()
            nb[:, :w - off] = f32(buf[:, :w - off] + buf[:, off:]).double()
            buf = nb
            off *= 2
        return f32(buf[:, 0])
    def tree_desc(vals, n):           # offsets w/2, w/4, ...
        w = 1
        while w < n: w *= 2
        buf = torch.zeros(Bn, w, device="cuda", dtype=torch.float64); buf[:, :n] = vals
        off = w // 2
        while off >= 1:
            nb = buf.clone()
            nb[:, :off] = f32(buf[:, :off] + buf[:, off:2 * off]).double()
            buf = nb
            off //= 2
        return f32(buf[:, 0])
    def acc4(vals, n, combine):
        accs = [torch.zeros(Bn, device="cuda") for _ in range(4)]
        for j in range(n):
            accs[j % 4] = f32(accs[j % 4].double() + vals[:, j])
        if combine == "seq":
            r = accs[0]
            for t in accs[1:]:
                r = f32(r.double() + t.double())
            return r
        return f32(f32(accs[0].double() + accs[1].double()).double() + f32(accs[2].double() + accs[3].double()).double())
    c = {"seq": seq, "tree_asc": tree(vd, ke), "tree_desc": tree_desc(vd, ke), "acc4_seq": acc4(vd, ke, "seq"), "acc4_pair": acc4(vd, ke, "pair")}
    print("   masked sum:", {k: round(float((vv == tot).float().mean()), 5) for k, vv in c.items()})
    # and the CPU for comparison
    tot_cpu = torch.sum(v.cpu(), -1).cuda()
    print("   cpu sum == seq:", float((tot_cpu == seq).float().mean()), " cpu==cuda:", float((tot_cpu == tot).float().mean()))
    mm_cpu = torch.matmul(src.cpu(), dst.cpu().permute(0, 2, 1)).squeeze(1).cuda()
    print("   cpu matmul == sep:", float((mm_cpu == cands["sep (x+y)+z"]).float().mean()), " == fma asc:", float((mm_cpu == cands["fma asc"]).float().mean()))
    e = torch.full((Bn,), 2.718281828459045, device="cuda")
    x = (torch.rand(Bn, device="cuda") - 0.5) * 40
    pw = e.pow(x)
    print("   pow == exp(x*log(e32))?", float((pw == torch.exp(x * torch.log(e))).float().mean()), " cpu pow == cuda pow:", float((e.cpu().pow(x.cpu()).cuda() == pw).float().mean()))
