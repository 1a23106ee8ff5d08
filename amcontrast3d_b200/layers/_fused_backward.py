"""Backward of the fused  grouping -> conv -> BatchNorm -> ReLU -> max  operator (layers/fused.py).

With  yhat = (y - mean) * invstd,  out = max_s relu(gamma * yhat + beta)  and  G = dL/dout, the gradient w.r.t.
the conv output is
    dL/dy[p,o] = ghat_o * D[p,o]  -  c0_o  -  c1_o * y[p,o]
    D[p,o] = G[q,o] [out > 0] [p is the arg-max of (q,o)],   ghat = gamma * invstd,
    c1 = ghat * invstd * dgamma / P,   c0 = ghat * dbeta / P - c1 * mean,   dbeta = sum D,   dgamma = sum D yhat.
The D term has ONE non-zero per (query, channel): it is handled by the two device kernels of
amc3d_fused_sa_backward_sparse.  The other two terms touch every grouped position, but they are linear in
x[p] = [f[idx[p]] | dp[p]], so their contributions to dW and df only need  Sx = sum_p x  and  Sxx = sum_p x x^T,
and because the feature part of x[p] is a row of f these are per-support-point sums weighted by
cnt[n] = #{p: idx[p] = n}:   sum_p f f^T = sum_n cnt[n] f[n] f[n]^T   — (B*N) x C x C GEMMs, 1/nsample of the
convolution's work, done here with torch.matmul (plain library GEMMs):
    dW'  =  E  -  c0 (x) Sx  -  diag(c1) W' Sxx
    df[n] = (sparse scatter)  -  cnt[n] v_f  -  Q_ff (cnt[n] f[n])  -  Q_fd dpsum[n],   Q = W'^T diag(c1) W',  v = W'^T c0
W' is the conv weight in the packed column order [features | dp] of fused.pack_weight.
"""
from __future__ import annotations

import torch

from .. import _capi
from .._capi import ptr, stream

_PREC = {"tf32": 1, "tf32x3": 3}


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
    gy = torch.empty((B * M, O), dtype=torch.float32, device=dev)
    dbg = torch.empty((2 * O,), dtype=torch.float64, device=dev)
    cnt = torch.empty((B, N), dtype=torch.float32, device=dev)
    dpsum = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
    mom = torch.empty((12,), dtype=torch.float64, device=dev)
    with _capi.guard(fT):
        st = stream(fT)
        _capi.call("amc3d_fused_sa_backward_prep", B, M, O, ptr(G), ptr(out), ptr(ysel), ptr(mean), ptr(invstd),
                   ptr(gamma), ptr(gy), ptr(dbg), st)
        _capi.call("amc3d_fused_sa_moments", B, N, M, ns, radius, int(normalize_dp), ptr(p), ptr(q), ptr(idx), ptr(cnt),
                   ptr(dpsum), ptr(mom), st)
    # the FP32-faithful mode keeps the small dense algebra in FP64; the TF32 mode uses FP32 library GEMMs
    wd = torch.float64 if precision == "tf32x3" else torch.float32
    dbeta, dgamma = dbg[:O], dbg[O:]
    ghat = gamma.double() * invstd.double()
    c1 = ghat * invstd.double() * dgamma / P
    c0 = ghat * dbeta / P - c1 * mean.double()
    W = wp[:, :Kq].to(wd)                                    # (O, Kq) packed [f | dp]
    f2 = fT.reshape(-1, C).to(wd)
    Fw = f2 * cnt.reshape(-1, 1).to(wd)                      # cnt[n] f[n]
    dps = dpsum.reshape(-1, 3).to(wd)
    Sff = f2.t() @ Fw                                        # sum_n cnt f f^T
    Sfd = f2.t() @ dps                                       # sum_p f dp^T
    Sdd = mom[3:12].view(3, 3).to(wd)
    Sxx = torch.cat([torch.cat([Sff, Sfd], 1), torch.cat([Sfd.t(), Sdd], 1)], 0)          # (Kq, Kq)
    Sx = torch.cat([Fw.sum(0), mom[0:3].to(wd)])
    c0w, c1w = c0.to(wd), c1.to(wd)
    dW_dense = -(c0w[:, None] * Sx[None, :]) - c1w[:, None] * (W @ Sxx)                   # (O, Kq)
    Qm = W.t() @ (c1w[:, None] * W)                                                       # (Kq, Kq)
    v = W.t() @ c0w
    dfT = -(cnt.reshape(-1, 1).to(wd) * v[None, :C]) - Fw @ Qm[:C, :C].t() - dps @ Qm[:C, C:].t()
    dfT = dfT.to(torch.float32).contiguous()                 # (B*N, C): the sparse scatter adds into it
    Op = (O + 31) // 32 * 32
    wT = torch.zeros((C, Op), dtype=torch.float32, device=dev)
    wT[:, :O] = wp[:, :C].t()
    E = torch.zeros((O, C + 8), dtype=torch.float32, device=dev)
    with _capi.guard(fT):
        _capi.call("amc3d_fused_sa_backward_sparse", B, N, M, C, O, Op, ns, radius, int(normalize_dp), _PREC[precision],
                   ptr(fT), ptr(p), ptr(q), ptr(idx), ptr(gy), ptr(arg), ptr(wT), ptr(dfT), ptr(E), stream(fT))
    dWp = E[:, :Kq].to(wd) + dW_dense
    dW = torch.cat([dWp[:, C:], dWp[:, :C]], 1).to(torch.float32).reshape(wshape)
    df = torch.empty((B, C, N), dtype=torch.float32, device=dev)
    with _capi.guard(fT):
        _capi.call("amc3d_transpose_batched", B, N, C, ptr(dfT), ptr(df), stream(fT))
    return (df, dW, dgamma.to(torch.float32), dbeta.to(torch.float32), None, None, None, None, None, None, None)
