"""Backward of the fused  grouping -> conv -> BatchNorm -> ReLU -> max  operator (layers/fused.py).

With  yhat = (y - mean) * invstd,  out = max_s relu(gamma * yhat + beta)  and  G = dL/dout, the gradient w.r.t.
the conv output is
    dL/dy[p,o] = ghat_o * D[p,o]  -  c0_o  -  c1_o * y[p,o]
    D[p,o] = G[q,o] [out > 0] [p is the arg-max of (q,o)],   ghat = gamma * invstd,
    c1 = ghat * invstd * dgamma / P,   c0 = ghat * dbeta / P - c1 * mean,   dbeta = sum D,   dgamma = sum D yhat.

* The D term has ONE non-zero per (query, channel), and where it lands only depends on which support point n the
  arg-max position refers to.  amc3d_fused_sa_backward_scatter adds ghat_o G'[q,o] into A[n, o]; then
      dW_f = A^T f ,    df = A W_f          ((B*N) x O x C GEMMs, 1/nsample of the convolution's work)
  and the three relative-coordinate columns of dW come out of the same kernel.
* The other two terms touch every grouped position, but they are linear in x[p] = [f[idx[p]] | dp[p]], so they only
  need  Sx = sum_p x  and  Sxx = sum_p x x^T,  which are per-support-point sums weighted by
  cnt[n] = #{p: idx[p] = n}  (amc3d_fused_sa_moments):   sum_p f f^T = sum_n cnt[n] f[n] f[n]^T :
      dW'  -=  c0 (x) Sx  +  diag(c1) W' Sxx
      df[n] -=  cnt[n] v_f  +  Q_ff (cnt[n] f[n])  +  Q_fd dpsum[n],      Q = W'^T diag(c1) W',  v = W'^T c0
All products here are plain library GEMMs (torch.matmul).  The per-support-point quantities are kept as the column
blocks of ONE matrix  X = [A | cnt*f | dpsum | cnt]  of shape (B*N, O+C+4) that the two kernels fill in place, so that
    f^T X  = [ (A^T f)^T | sum_n cnt f f^T | sum_p f dp^T | sum_n cnt f ]     (every moment dW needs, one GEMM)
    df     = X [ W_f ; -Q_ff^T ; -Q_fd^T ; -v_f^T ]                             (one GEMM)
instead of eight products and their read-modify-write accumulations over (B*N)-row arrays.  W' is the conv weight in
the packed column order [features | dp] of fused.pack_weight.
"""
from __future__ import annotations

import torch

from .. import _capi
from .._capi import ptr, stream


def _tall_tn(a, b, B, N):
    """a^T b for two tall matrices with B*N rows: split the long reduction into row blocks (batched GEMM + a sum of the
    partial products) — the library's single-GEMM heuristics launch a handful of CTAs for a (C x B*N) @ (B*N x K) product"""
    s = next((d for d in (16, 12, 10, 8, 6, 5, 4, 3, 2) if N % d == 0 and N // d >= 1024), 1)
    nb = B * s
    if nb == 1:
        return a.t() @ b
    part = torch.bmm(a.view(nb, -1, a.shape[1]).transpose(1, 2), b.view(nb, -1, b.shape[1]))
    return part.sum(0)


def backward(ctx, grad_out):
    fT, wp, gamma, q, p, idx, ysel, arg, mean, invstd, out = ctx.saved_tensors
    radius, normalize_dp, precision, wshape = ctx.cfg
    B, N, C = fT.shape
    M, ns = idx.shape[1], idx.shape[2]
    O = wp.shape[0]
    Kq = C + 3
    P = float(B * M * ns)
    dev = fT.device
    G = grad_out.contiguous().float()
    # ONE (B*N, O + C + 4) matrix  X = [ A | cnt * f | dpsum | cnt ]: the kernels write their column blocks in place
    # (row stride Kc), so that the feature gradient is one GEMM with it and every weight-gradient moment another
    Kc = O + C + 4
    X = torch.empty((B * N, Kc), dtype=torch.float32, device=dev)
    A, Fw, dps, cnt = X[:, :O], X[:, O:O + C], X[:, O + C:O + C + 3], X[:, O + C + 3:]
    red = torch.empty((5 * O,), dtype=torch.float64, device=dev)
    mom = torch.empty((12,), dtype=torch.float64, device=dev)
    with _capi.guard(fT):
        st = stream(fT)
        _capi.call("amc3d_fused_sa_backward_scatter", B, N, M, O, ns, radius, int(normalize_dp), ptr(G), ptr(out),
                   ptr(ysel), ptr(arg), ptr(idx), ptr(p), ptr(q), ptr(mean), ptr(invstd), ptr(gamma), ptr(A), Kc, ptr(red), st)
        _capi.call("amc3d_fused_sa_moments", B, N, M, ns, radius, int(normalize_dp), ptr(p), ptr(q), ptr(idx), ptr(cnt), Kc,
                   ptr(dps), Kc, ptr(mom), st)
    # O-sized coefficient vectors in FP64 (amc3d_fused_sa_backward_coefs); the GEMMs in FP32 (TF32 tensor cores in
    # 'tf32' mode, as the reference's cuDNN backward would use; exact FP32 in the FP32-faithful mode)
    f2 = fT.view(B * N, C)
    Kp = wp.shape[1]                                         # C + 8: small matrices padded like W' (GEMM dims % 8 == 0)
    c1wx = torch.empty((O, Kp), dtype=torch.float32, device=dev)              # [ c1 (.) W' | c0 | 0 ]
    dgamma = torch.empty((O,), dtype=torch.float32, device=dev)
    dbeta = torch.empty((O,), dtype=torch.float32, device=dev)
    sxx = torch.empty((Kp, Kp), dtype=torch.float32, device=dev)
    wc = torch.empty((Kc, C), dtype=torch.float32, device=dev)
    dWp = torch.empty((O, Kp), dtype=torch.float32, device=dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    try:
        with _capi.guard(fT):
            _capi.call("amc3d_fused_sa_backward_coefs", C, O, P, ptr(gamma), ptr(invstd), ptr(mean), ptr(red), ptr(wp),
                       ptr(c1wx), ptr(dgamma), ptr(dbeta), st)
        torch.mul(f2, cnt, out=Fw)                           # cnt[n] f[n]
        G1 = _tall_tn(f2, X, B, N)                           # (C, Kc) = [ (A^T f)^T | sum cnt f f^T | sum_p f dp^T | sum cnt f ]
        QV = wp.t() @ c1wx                                   # (Kp, Kp) = [ W'^T diag(c1) W' | W'^T c0 | 0 ]
        with _capi.guard(fT):
            _capi.call("amc3d_fused_sa_backward_assemble", C, O, ptr(G1), ptr(mom), ptr(red), ptr(wp), ptr(QV), ptr(c1wx),
                       ptr(sxx), ptr(wc), ptr(dWp), st)
        dWp.addmm_(c1wx, sxx, alpha=-1.0)                    # - diag(c1) W' S_xx  (the c0 column meets a zero row)
        dfT = X @ wc                                         # A W_f - (cnt f) Q_ff - dpsum Q_fd^T - cnt v_f     (B*N, C)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    dW = torch.cat([dWp[:, C:Kq], dWp[:, :C]], 1).reshape(wshape)
    df = torch.empty((B, C, N), dtype=torch.float32, device=dev)
    with _capi.guard(fT):
        _capi.call("amc3d_transpose_batched", B, N, C, ptr(dfT), ptr(df), stream(fT))
    return (df, dW, dgamma, dbeta, None, None, None, None, None, None, None)
