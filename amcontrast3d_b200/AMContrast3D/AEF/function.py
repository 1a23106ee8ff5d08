"""Mirror of openpoints/AMContrast3D/AEF/function.py:8-39 (small torch helpers kept for API parity;
the fused ambiguity kernel evaluates the same expressions in the same FP32 order on device)."""
import torch

_inf = 1e9
_eps = 1e-12


def inverse_sigmoid_function(cc, t, b):
    """a = 1 / (1 + t ** (b * cc)), t = e (function.py:10-14)"""
    return 1 / (1 + t.pow(b * cc))


def square_distance(src, dst):
    """|src|^2 + |dst|^2 - 2 src.dst^T for src (B,N,C), dst (B,M,C) -> (B,N,M) (function.py:18-39)"""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist
