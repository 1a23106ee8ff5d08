"""The fused  grouping -> 1x1 conv -> BatchNorm (training) -> ReLU -> max  operator (csrc/fused_sa.cu,
layers/fused.py) on the GPU against
  (a) tests/golden/fused_golden.npz — outputs, gradients and running statistics of the REFERENCE's own
      LocalAggregation / SetAbstraction modules (pointnext_AA.py:20-63, 78-166) run on CPU
      (tests/golden/make_fused_golden.py), within 2e-5 relative L2 in the FP32-faithful mode (3 x TF32);
  (b) the oracle restatement (oracle/fused_oracle.py) at PointNeXt-XL layer shapes, same bar;
  (c) in TF32 mode — the precision of the reference's cuDNN convolution under torch's default
      `cudnn.allow_tf32 = True` — within 2e-3 (10-bit mantissa operands).
The modules are the reference's own classes when its Python is available (oracle/ref_python.py), built by its
own constructors, with `forward` routed through layers.fused — the drop-in proof for this operator."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import REPO, rel_err

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
pytestmark = pytest.mark.gpu
PATH = os.path.join(REPO, "tests", "golden", "fused_golden.npz")
TOL = 2e-5
TOL_TF32 = 2e-3


class _Attr(dict):
    """dict with attribute access (what the reference's EasyDict config nodes offer)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


def _module(kind, cin, cout, stride, radius, nsample):
    """the reference's module when importable, else a structural stand-in with the same attributes"""
    from oracle import ref_python as rp
    import torch.nn as nn
    from amcontrast3d_b200.layers import QueryAndGroup
    if rp.available():
        rp.import_reference(1)
        from openpoints.models.backbone.pointnext_AA import LocalAggregation, SetAbstraction
        ga = _Attr(NAME="ballquery", radius=radius, nsample=nsample, normalize_dp=True)
        common = dict(norm_args={"norm": "bn"}, act_args={"act": "relu"}, conv_args={"order": "conv-norm-act"})
        if kind == "la":
            return LocalAggregation([cin, cout], group_args=ga, feature_type="dp_fj", reduction="max", **common)
        return SetAbstraction(cin, cout, layers=1, stride=stride, group_args=ga, feature_type="dp_fj", use_res=False, **common)
    m = nn.Module()
    m.convs = nn.Sequential(nn.Sequential(nn.Conv2d(cin + 3, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU()))
    m.grouper = QueryAndGroup(radius, nsample, normalize_dp=True)
    m.feature_type, m.reduction, m.stride = "dp_fj", "max", stride
    m.is_head, m.all_aggr, m.use_res = False, False, False
    return m


@pytest.mark.parametrize("name", ["la_c32", "la_c64_ns32", "sa_c32_c64"])
def test_fused_operator_vs_reference_golden(name):
    from make_fused_golden import CASES, fused_inputs
    from amcontrast3d_b200.layers import fused
    g = np.load(PATH)
    kind, B, N, cin, cout, stride, radius, nsample = CASES[name]
    inp = fused_inputs(name)
    mod = _module(kind, cin, cout, stride, radius, nsample).cuda()
    conv, bn = mod.convs[0][0], mod.convs[0][1]
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(inp["w"]).view(cout, cin + 3, 1, 1))
        bn.weight.copy_(torch.from_numpy(inp["gamma"]))
        bn.bias.copy_(torch.from_numpy(inp["beta"]))
    mod.train()
    p = torch.from_numpy(inp["xyz"]).cuda()
    f = torch.from_numpy(inp["f"]).cuda().requires_grad_(True)
    if kind == "la":
        y = fused.local_aggregation_forward(mod, (p, f), precision="tf32x3")
    else:
        new_p, y = fused.set_abstraction_forward(mod, (p, f), precision="tf32x3")
        assert np.array_equal(new_p.cpu().numpy(), g[f"{name}/new_p"])
    assert rel_err(y.detach().cpu().numpy(), g[f"{name}/y"]) < TOL
    y.backward(torch.from_numpy(inp["go"]).cuda())
    assert rel_err(f.grad.cpu().numpy(), g[f"{name}/grad_f"]) < TOL
    assert rel_err(conv.weight.grad.view(cout, cin + 3).cpu().numpy(), g[f"{name}/grad_w"]) < TOL
    assert rel_err(bn.weight.grad.cpu().numpy(), g[f"{name}/grad_gamma"]) < TOL
    assert rel_err(bn.bias.grad.cpu().numpy(), g[f"{name}/grad_beta"]) < TOL
    assert rel_err(bn.running_mean.cpu().numpy(), g[f"{name}/running_mean"]) < TOL
    assert rel_err(bn.running_var.cpu().numpy(), g[f"{name}/running_var"]) < TOL
    assert int(bn.num_batches_tracked) == 1


def _torch_composition(q, p, f, idx, w, gamma, beta, radius, arg, eps=1e-5, normalize_dp=True):
    """the module composition in FP64 on the GPU (gather + einsum + batch statistics + relu + max).  The max is
    taken at the sample `arg` (B,O,M) the operator chose; the second return value is how far below the true
    maximum that choice is.  Max-pooling over values that carry rounding errors has no unique arg-max: two
    samples whose conv outputs agree to 1e-7 are interchangeable for the forward, but route the gradient to
    different neighbours, so gradients are compared on the operator's own (verified) choice."""
    B, C, N = f.shape
    M, ns = idx.shape[1], idx.shape[2]
    flat = idx.reshape(B, 1, -1).long()
    dp = torch.gather(p.transpose(1, 2), 2, flat.expand(-1, 3, -1)).reshape(B, 3, M, ns)
    scale = (1.0 / np.float32(radius)) if normalize_dp else 1.0
    dp = ((dp - q.transpose(1, 2).unsqueeze(-1)) * scale).to(f.dtype)     # FP32 values, as the operator forms them
    fj = torch.gather(f, 2, flat.expand(-1, C, -1)).reshape(B, C, M, ns)
    y = torch.einsum("oc,bcps->bops", w, torch.cat([dp, fj], 1))
    mean = y.mean(dim=(0, 2, 3), keepdim=True)
    var = y.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    z = torch.relu((y - mean) / torch.sqrt(var + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1))
    chosen = torch.gather(z, -1, arg.long().unsqueeze(-1)).squeeze(-1)
    return chosen, (z.max(-1)[0] - chosen).detach().abs().max()


@pytest.mark.parametrize("B,N,M,C,O,ns,radius", [
    (2, 4096, 1024, 64, 128, 32, 0.1),      # SetAbstraction, level 1 of PointNeXt-XL (C 64 -> 128, stride 4)
    (8, 6000, 6000, 128, 128, 32, 0.2),     # LocalAggregation, level 1 at BASELINE config 2 size (1.5 M grouped positions)
    (2, 375, 375, 512, 512, 32, 0.8),       # level 3: four output-channel slices per position tile
    (3, 93, 93, 1024, 1024, 32, 1.6),       # level 4; 3 * 93 queries: a ragged last tile
    (2, 500, 500, 40, 72, 16, 0.3),         # nsample 16, channel counts off the 32 / 128 grids
    (2, 300, 300, 264, 72, 32, 0.3),        # one output slice whose weights do not stay resident (C > 216)
])
def test_fused_operator_vs_fp64_composition(B, N, M, C, O, ns, radius):
    from amcontrast3d_b200 import scenes
    from amcontrast3d_b200.layers import ball_query, furthest_point_sample
    from amcontrast3d_b200.layers.fused import FusedGroupConvBNReLUMax
    xyz, _ = scenes.batch_of_scenes(B, N, "surface", first_scene=31)
    p = torch.from_numpy(xyz).cuda()
    q = p
    if M != N:
        i = furthest_point_sample(p, M).long()
        q = torch.gather(p, 1, i.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    g = torch.Generator(device="cuda").manual_seed(17)
    f = torch.randn(B, C, N, device="cuda", generator=g)
    w = torch.randn(O, C + 3, device="cuda", generator=g) / (C + 3) ** 0.5
    gamma = 1 + 0.1 * torch.randn(O, device="cuda", generator=g)
    gamma[::7] *= -1                                                     # negative BatchNorm weights: the min branch
    beta = 0.1 * torch.randn(O, device="cuda", generator=g)
    go = torch.randn(B, O, M, device="cuda", generator=g)
    idx = ball_query(radius, ns, p, q)
    # gap: how far below the true maximum the chosen sample may be.  FP32-faithful mode: rounding only;
    # TF32 operands (what the reference's cuDNN convolution reads under torch's default): 10-bit mantissas
    for prec, tol, gap in (("tf32x3", TOL, 2e-5), ("tf32", TOL_TF32, 2e-2)):
        mine = [t.clone().requires_grad_(True) for t in (f, w, gamma, beta)]
        out, mean, var = FusedGroupConvBNReLUMax.apply(*mine, q, p, idx, radius, True, 1e-5, prec)
        arg = out.grad_fn.saved_tensors[7].view(B, M, O).transpose(1, 2).clone() # (B,O,M) u8, the operator's arg-max
        out.backward(go)
        ref_leaves = [t.double().requires_grad_(True) for t in (f, w, gamma, beta)]
        ref_out, below = _torch_composition(q.double(), p.double(), ref_leaves[0], idx, *ref_leaves[1:], radius, arg)
        ref_out.backward(go.double())
        assert float(below) <= gap, (prec, float(below))
        assert rel_err(out.detach().cpu().numpy(), ref_out.detach().cpu().numpy()) < tol, prec
        # gradients: TF32 operands in the dX product as well, and sums of ~1e6 TF32-rounded terms in dW
        gtol = tol if prec == "tf32x3" else 5e-3
        for a, b, what in zip(mine, ref_leaves, ("df", "dW", "dgamma", "dbeta")):
            assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < gtol, (prec, what)


def test_fused_operator_without_normalised_relative_coordinates_and_single_query():
    """normalize_dp = False (QueryAndGroup's other setting) and the smallest launch there is: one query"""
    from amcontrast3d_b200 import scenes
    from amcontrast3d_b200.layers import ball_query
    from amcontrast3d_b200.layers.fused import FusedGroupConvBNReLUMax
    for B, N, M, C, O, ns, radius in ((2, 400, 400, 32, 48, 32, 0.3), (1, 64, 1, 16, 8, 16, 0.5)):
        xyz, _ = scenes.batch_of_scenes(B, N, "surface", first_scene=11)
        p = torch.from_numpy(xyz).cuda()
        q = p[:, :M].contiguous()
        g = torch.Generator(device="cuda").manual_seed(23)
        f = torch.randn(B, C, N, device="cuda", generator=g)
        w = torch.randn(O, C + 3, device="cuda", generator=g) / (C + 3) ** 0.5
        gamma = 1 + 0.1 * torch.randn(O, device="cuda", generator=g)
        gamma[::3] *= -1
        beta = 0.1 * torch.randn(O, device="cuda", generator=g)
        go = torch.randn(B, O, M, device="cuda", generator=g)
        idx = ball_query(radius, ns, p, q)
        mine = [t.clone().requires_grad_(True) for t in (f, w, gamma, beta)]
        out, mean, var = FusedGroupConvBNReLUMax.apply(*mine, q, p, idx, radius, False, 1e-5, "tf32x3")
        arg = out.grad_fn.saved_tensors[7].view(B, M, O).transpose(1, 2).clone()
        out.backward(go)
        ref_leaves = [t.double().requires_grad_(True) for t in (f, w, gamma, beta)]
        ref_out, below = _torch_composition(q.double(), p.double(), ref_leaves[0], idx, *ref_leaves[1:], radius, arg,
                                            normalize_dp=False)
        ref_out.backward(go.double())
        assert float(below) <= 2e-5
        assert rel_err(out.detach().cpu().numpy(), ref_out.detach().cpu().numpy()) < TOL
        for a, b, what in zip(mine, ref_leaves, ("df", "dW", "dgamma", "dbeta")):
            assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < TOL, what


def test_unsupported_configurations_use_the_module_composition():
    """eval mode / several conv layers are not fused: the entry points report it instead of guessing"""
    from amcontrast3d_b200.layers import fused
    mod = _module("la", 32, 32, 1, 0.2, 16).cuda()
    mod.eval()
    p = torch.rand(1, 64, 3, device="cuda")
    f = torch.randn(1, 32, 64, device="cuda")
    assert fused._fusable(mod, f) is None
    assert not fused.supported(30, 32) and not fused.supported(32, 24) and fused.supported(64, 32)


def test_bound_reference_module_runs_the_fused_operator():
    """what compat.install(tier=4) does to the reference's classes (layers/fused.bind), on a subclass so that the
    other tests of this process keep the unbound reference: the module call a PointNeXt backbone makes runs the
    fused operator, and a configuration it does not cover falls back to the reference's composition"""
    from oracle import ref_python as rp
    if not rp.available():
        pytest.skip("reference Python not available")
    rp.import_reference(1)
    from amcontrast3d_b200.layers import fused
    from make_fused_golden import CASES, fused_inputs
    g = np.load(PATH)
    name = "la_c64_ns32"
    kind, B, N, cin, cout, stride, radius, nsample = CASES[name]
    inp = fused_inputs(name)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                 # -> the FP32-faithful mode is the default precision
    try:
        ref_mod = _module(kind, cin, cout, stride, radius, nsample)
        Bound = fused.bind(type("BoundLocalAggregation", (type(ref_mod),), {}), "la")
        assert Bound._amc3d_fused and not getattr(type(ref_mod), "_amc3d_fused", False)
        ref_mod.__class__ = Bound
        mod = ref_mod.cuda()
        conv, bn = mod.convs[0][0], mod.convs[0][1]
        with torch.no_grad():
            conv.weight.copy_(torch.from_numpy(inp["w"]).view(cout, cin + 3, 1, 1))
            bn.weight.copy_(torch.from_numpy(inp["gamma"]))
            bn.bias.copy_(torch.from_numpy(inp["beta"]))
        mod.train()
        p = torch.from_numpy(inp["xyz"]).cuda()
        f = torch.from_numpy(inp["f"]).cuda()
        y = mod((p, f))
        assert rel_err(y.detach().cpu().numpy(), g[f"{name}/y"]) < TOL
        assert int(bn.num_batches_tracked) == 1             # the fused path updated the running statistics
        mod.eval()                                          # not covered: the reference's own composition answers
        y_eval = mod((p, f))
        assert y_eval.shape == y.shape and torch.isfinite(y_eval).all()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
