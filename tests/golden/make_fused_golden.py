"""Generate tests/golden/fused_golden.npz: outputs and gradients of the REFERENCE's own LocalAggregation and
SetAbstraction modules (openpoints/models/backbone/pointnext_AA.py:20-63, 78-166) in training mode — the
parity target of the fused  grouping -> 1x1 conv -> BatchNorm -> ReLU -> max  operator (SURVEY.md §8f rank 1,
DESIGN.md §8).

Run in the build container only (needs /root/reference):   python tests/golden/make_fused_golden.py
The modules are imported unmodified and run on CPU with the stubs of make_loss_golden.py plus a fake
`pointnet2_batch_cuda` whose wrappers are backed by oracle/ops_oracle.c (the literal restatements of the
reference kernels).  Module parameters and inputs are regenerated from seeds by fused_inputs(); the file
holds the reference's outputs only."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_loss_golden as mlg  # noqa: E402
from oracle import ops_oracle as oo  # noqa: E402

CASES = {
    # name: (kind, B, N, C_in, C_out, stride, radius, nsample)
    "la_c32": ("la", 2, 600, 32, 32, 1, 0.2, 16),
    "la_c64_ns32": ("la", 2, 400, 64, 64, 1, 0.25, 32),
    "sa_c32_c64": ("sa", 2, 800, 32, 64, 4, 0.15, 16),
}


def fused_inputs(name):
    kind, B, N, cin, cout, stride, radius, nsample = CASES[name]
    xyz, _ = mlg.scenes.batch_of_scenes(B, N, "surface", first_scene=31)
    rng = np.random.default_rng(17)
    f = rng.standard_normal((B, cin, N)).astype(np.float32)
    w = (rng.standard_normal((cout, cin + 3)) / np.sqrt(cin + 3)).astype(np.float32)     # 1x1 conv, no bias (BN follows)
    gamma = (1.0 + 0.1 * rng.standard_normal(cout)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    npoint = N // stride
    go = rng.standard_normal((B, cout, npoint)).astype(np.float32)                        # upstream gradient
    return dict(xyz=np.ascontiguousarray(xyz), f=f, w=w, gamma=gamma, beta=beta, go=go)


def install_fake_pointnet2():
    """pointnet2_batch_cuda on CPU tensors through the oracle (pointnet2_api.cpp:10-24 argument orders)"""
    m = types.ModuleType("pointnet2_batch_cuda")

    def ball_query_wrapper(b, n, npoint, radius, nsample, new_xyz, xyz, idx):
        idx.copy_(torch.from_numpy(oo.ball_query(radius, nsample, xyz.numpy(), new_xyz.numpy())))

    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        out.copy_(torch.from_numpy(oo.group_points(points.detach().numpy(), idx.numpy())))

    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        grad_points.copy_(torch.from_numpy(oo.group_points_grad(grad_out.numpy(), idx.numpy(), n)))

    def furthest_point_sampling_wrapper(b, n, m_, xyz, temp, idx):
        i, _ = oo.fps(xyz.numpy(), m_)
        idx.copy_(torch.from_numpy(i))

    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        out.copy_(torch.from_numpy(oo.gather_points(points.detach().numpy(), idx.numpy())))

    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        grad_points.copy_(torch.from_numpy(oo.gather_points_grad(grad_out.numpy(), idx.numpy(), n)))

    for fn in (ball_query_wrapper, group_points_wrapper, group_points_grad_wrapper, furthest_point_sampling_wrapper,
               gather_points_wrapper, gather_points_grad_wrapper):
        setattr(m, fn.__name__, fn)
    sys.modules["pointnet2_batch_cuda"] = m


class AttrDict(dict):
    """dict with attribute access (what the reference's EasyDict config nodes offer)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


def main():
    mlg.install_stubs()
    install_fake_pointnet2()
    import openpoints.models  # noqa: F401
    from openpoints.models.backbone.pointnext_AA import LocalAggregation, SetAbstraction

    out = {}
    for name, (kind, B, N, cin, cout, stride, radius, nsample) in CASES.items():
        inp = fused_inputs(name)
        group_args = AttrDict(NAME="ballquery", radius=radius, nsample=nsample, normalize_dp=True)
        common = dict(norm_args={"norm": "bn"}, act_args={"act": "relu"}, conv_args={"order": "conv-norm-act"})
        if kind == "la":
            mod = LocalAggregation([cin, cout], group_args=group_args, feature_type="dp_fj", reduction="max", **common)
        else:
            mod = SetAbstraction(cin, cout, layers=1, stride=stride, group_args=group_args, feature_type="dp_fj",
                                 use_res=False, **common)
        conv, bn = mod.convs[0][0], mod.convs[0][1]
        assert conv.weight.shape == (cout, cin + 3, 1, 1) and conv.bias is None and isinstance(bn, torch.nn.BatchNorm2d)
        with torch.no_grad():
            conv.weight.copy_(torch.from_numpy(inp["w"]).view(cout, cin + 3, 1, 1))
            bn.weight.copy_(torch.from_numpy(inp["gamma"]))
            bn.bias.copy_(torch.from_numpy(inp["beta"]))
        mod.train()
        p = torch.from_numpy(inp["xyz"])
        f = torch.from_numpy(inp["f"]).requires_grad_(True)
        if kind == "la":
            y = mod((p, f))
        else:
            new_p, y = mod((p, f))
            out[f"{name}/new_p"] = new_p.numpy()
        y.backward(torch.from_numpy(inp["go"]))
        out[f"{name}/y"] = y.detach().numpy()
        out[f"{name}/grad_f"] = f.grad.numpy()
        out[f"{name}/grad_w"] = conv.weight.grad.view(cout, cin + 3).numpy()
        out[f"{name}/grad_gamma"] = bn.weight.grad.numpy()
        out[f"{name}/grad_beta"] = bn.bias.grad.numpy()
        out[f"{name}/running_mean"] = bn.running_mean.numpy()
        out[f"{name}/running_var"] = bn.running_var.numpy()
        out[f"{name}/bn_eps_momentum"] = np.array([bn.eps, bn.momentum])
        print(name, "y", tuple(y.shape), "mean", float(y.mean()), "|grad_f|", float(f.grad.norm()))
    path = os.path.join(HERE, "fused_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
