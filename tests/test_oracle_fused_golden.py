"""oracle/fused_oracle.py (grouping -> 1x1 conv -> training-mode BatchNorm -> ReLU -> max) against outputs and
gradients of the REFERENCE's own LocalAggregation / SetAbstraction modules run on CPU
(tests/golden/fused_golden.npz, tests/golden/make_fused_golden.py).  This pins the parity target of the fused
operator planned next; tolerance 2e-5 relative L2 (einsum vs Conv2d accumulation order)."""
import os
import sys

import numpy as np
import pytest
import torch

from _util import REPO, rel_err
from oracle import fused_oracle as fo

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
PATH = os.path.join(REPO, "tests", "golden", "fused_golden.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="fused_golden.npz not generated yet")
TOL = 2e-5


@pytest.mark.parametrize("name", ["la_c32", "la_c64_ns32", "sa_c32_c64"])
def test_fused_operator_oracle(name):
    from make_fused_golden import CASES, fused_inputs
    g = np.load(PATH)
    kind, B, N, cin, cout, stride, radius, nsample = CASES[name]
    inp = fused_inputs(name)
    xyz = torch.from_numpy(inp["xyz"])
    f = torch.from_numpy(inp["f"]).requires_grad_(True)
    w = torch.from_numpy(inp["w"]).requires_grad_(True)
    gamma = torch.from_numpy(inp["gamma"]).requires_grad_(True)
    beta = torch.from_numpy(inp["beta"]).requires_grad_(True)
    eps, momentum = g[f"{name}/bn_eps_momentum"]
    if kind == "la":
        y, mean, var = fo.local_aggregation(xyz, f, w, gamma, beta, radius, nsample, eps=float(eps))
    else:
        new_p, y, mean, var = fo.set_abstraction(xyz, f, w, gamma, beta, stride, radius, nsample, eps=float(eps))
        assert np.array_equal(new_p.numpy(), g[f"{name}/new_p"])
    assert rel_err(y.detach().numpy(), g[f"{name}/y"]) < TOL
    y.backward(torch.from_numpy(inp["go"]))
    assert rel_err(f.grad.numpy(), g[f"{name}/grad_f"]) < TOL
    assert rel_err(w.grad.numpy(), g[f"{name}/grad_w"]) < TOL
    assert rel_err(gamma.grad.numpy(), g[f"{name}/grad_gamma"]) < TOL
    assert rel_err(beta.grad.numpy(), g[f"{name}/grad_beta"]) < TOL
    # running statistics after one training step from (0, 1): momentum * batch statistic (unbiased variance)
    assert rel_err(float(momentum) * mean.detach().numpy(), g[f"{name}/running_mean"]) < TOL
    assert rel_err((1 - float(momentum)) + float(momentum) * var.detach().numpy(), g[f"{name}/running_var"]) < TOL
