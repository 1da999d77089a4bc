"""torch-fp64 twin of ``oracle/kalman_numpy.py`` (ORACLE - test infrastructure).

The reference's gradient is PyTensor reverse-mode autodiff of the scan graph in
``/root/reference/pymc_statespace/filters/kalman_filter.py`` (SURVEY.md section 8(a) row
a10); there is no hand-written reference gradient to restate.  This module writes
the same forward equations with torch ops on CPU float64 tensors so that
``torch.autograd`` provides the gradient oracle ("generic-op gauge": every entry of
every input matrix is an independent variable, each op differentiated as the
generic dense function it is - see DESIGN.md "gradient gauge").

DARE / Lyapunov are wrapped as autograd Functions whose backward is the formula
the reference itself uses (``utils/pytensor_scipy.py:39-60``; PyTensor's
``SolveDiscreteLyapunov`` grad).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import torch

from .kalman_numpy import FILTER_KINDS, LOG_2PI

DT = torch.float64


class _Lyapunov(torch.autograd.Function):
    """X = A X A^T + Q (scipy bilinear, as models/SARIMAX.py:106)."""

    @staticmethod
    def forward(ctx, A, Q):
        X = torch.from_numpy(
            scipy.linalg.solve_discrete_lyapunov(A.detach().numpy(), Q.detach().numpy(), method="bilinear")
        )
        ctx.save_for_backward(A, X)
        return X

    @staticmethod
    def backward(ctx, dX):
        A, X = ctx.saved_tensors
        S = torch.from_numpy(
            scipy.linalg.solve_discrete_lyapunov(A.numpy().T.copy(), dX.numpy().copy(), method="bilinear")
        )
        A_bar = S @ A @ X.T + S.T @ A @ X
        return A_bar, S


class _DARE(torch.autograd.Function):
    """SolveDiscreteARE, utils/pytensor_scipy.py:11-60."""

    @staticmethod
    def forward(ctx, A, B, Q, R):
        X = torch.from_numpy(
            scipy.linalg.solve_discrete_are(A.detach().numpy(), B.detach().numpy(), Q.detach().numpy(), R.detach().numpy())
        )
        ctx.save_for_backward(A, B, Q, R, X)
        return X

    @staticmethod
    def backward(ctx, dX):
        A, B, Q, R, X = ctx.saved_tensors
        K_inner = R + B.T @ X @ B
        K = torch.linalg.solve(K_inner, torch.eye(R.shape[0], dtype=DT)) @ B.T @ X @ A
        A_tilde = A - B @ K
        dX_symm = 0.5 * (dX + dX.T)
        S = torch.from_numpy(
            scipy.linalg.solve_discrete_lyapunov(A_tilde.numpy(), dX_symm.numpy(), method="bilinear")
        )
        A_bar = 2 * X @ A_tilde @ S
        B_bar = -2 * X @ A_tilde @ S @ K.T
        Q_bar = S
        R_bar = K @ S @ K.T
        return A_bar, B_bar, Q_bar, R_bar


solve_discrete_lyapunov = _Lyapunov.apply
solve_discrete_are = _DARE.apply


def _mask(y, Z, H):
    nan_mask = torch.isnan(y)
    all_nan = bool(nan_mask.all())
    W = torch.eye(y.shape[0], dtype=DT)
    idx = nan_mask.ravel()
    W[idx, idx] = 0.0
    return torch.where(nan_mask, torch.zeros_like(y), y), W @ Z, W @ H, all_nan


def _predict(a, P, c, T, R, Q):
    a_hat = T @ a + c
    P_hat = T @ P @ T.T + R @ Q @ R.T
    return a_hat, 0.5 * (P_hat + P_hat.T)


def _tri_lower_solve(A, B):
    # SolveTriangular(lower=True): only tril(A) is read (and only tril gets gradient)
    return torch.linalg.solve_triangular(torch.tril(A), B, upper=False)


def _update(kind, a, P, y, c, d, Z, H, flag, F_inv_ss, strict):
    m, p = P.shape[0], Z.shape[0]
    I_m, I_p = torch.eye(m, dtype=DT), torch.eye(p, dtype=DT)
    fl = 1.0 if flag else 0.0
    if kind == "standard":
        v = y - Z @ a - d
        PZT = P @ Z.T
        F = Z @ PZT + H
        F_inv = torch.linalg.solve(F + I_p * fl, I_p)
        K = PZT @ F_inv
        nconst = 1.0 if strict else float(p)
        ll = None if flag else -0.5 * (nconst * LOG_2PI + torch.log(torch.linalg.det(F)) + v.T @ F_inv @ v).ravel()[0]
    elif kind == "steady_state":
        v = y - Z @ a
        if not strict:
            v = v - d
        PZT = P @ Z.T
        F = Z @ PZT + H
        K = PZT @ F_inv_ss
        nconst = 1.0 if strict else float(p)
        ll = None if flag else -0.5 * (nconst * LOG_2PI + torch.log(torch.linalg.det(F)) + v.T @ F_inv_ss @ v).ravel()[0]
    elif kind == "cholesky":
        v = y - Z @ a - d
        PZT = P @ Z.T
        F = Z @ PZT + H + I_p * fl
        L = torch.linalg.cholesky(F)
        if strict:
            second = _tri_lower_solve
        else:
            second = lambda A, B: torch.linalg.solve_triangular(A, B, upper=True)  # noqa: E731
        K = second(L.T, _tri_lower_solve(L, PZT.T)).T * (1.0 - fl)
        inner = second(L.T, _tri_lower_solve(L, v))
        ll = None if flag else (-0.5 * (p * LOG_2PI + (v.T @ inner).ravel()) - torch.log(torch.diag(L)).sum()).ravel()[0]
    elif kind == "single":
        y_hat = (Z @ a).ravel() - d if strict else (Z @ a).ravel() + d
        v = y - y_hat
        PZT = P @ Z.T
        F = (Z @ PZT + H).ravel() + fl
        K = PZT / F
        ll = None if flag else (-0.5 * (LOG_2PI + torch.log(F) + v**2 / F)).ravel()[0]
    else:
        raise NotImplementedError(kind)
    I_KZ = I_m - K @ Z
    a_f = a + (K * v if kind == "single" else K @ v)
    P_f = I_KZ @ P @ I_KZ.T + K @ H @ K.T
    return a_f, P_f, ll


def _step_univariate(y, a, P, c, d, T, Z, R, H, Q):
    p = y.shape[0]
    nan_mask = torch.isnan(y).ravel()
    W = torch.eye(p, dtype=DT)
    W[nan_mask, nan_mask] = 0.0
    Zm, Hm = W @ Z, W @ H
    ym = torch.where(torch.isnan(y), torch.zeros_like(y), y)
    sigma = torch.diag(Hm)
    ll_sum, count = 0.0, 0
    for i in range(p):
        Zr = Zm[i][None, :]
        v = ym[i].reshape(1, 1) - Zr @ a - d[i]
        PZT = P @ Zr.T
        F = Zr @ PZT + sigma[i]
        flag = bool((F == 0).item()) or bool(nan_mask[i])
        if flag:
            continue  # every contribution is multiplied by (1 - flag) = 0
        K = PZT / F
        a = a + K * v
        P = P - torch.outer(K.ravel(), K.ravel()) * F
        lli = (torch.log(F) + v**2 / F).ravel()[0]
        if lli.item() != 0:
            count += 1
        ll_sum = ll_sum + lli
    a_hat, P_hat = _predict(a, P, c, T, R, Q)
    ll = -0.5 * (count * LOG_2PI + ll_sum)
    return a, a_hat, P, P_hat, ll


def kalman_filter(kind, data, a0, P0, T, Z, R, H, Q, c=None, d=None, strict_reference=True):
    """Same contract as ``kalman_numpy.kalman_filter`` but on torch tensors (grad-enabled)."""
    kind = kind.lower()
    assert kind in FILTER_KINDS
    as_t = lambda x: x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=DT)  # noqa: E731
    data, a0, P0, T, Z, R, H, Q = (as_t(x) for x in (data, a0, P0, T, Z, R, H, Q))
    n = data.shape[0]
    p, m = Z.shape[-2], Z.shape[-1]
    c = torch.zeros((m, 1), dtype=DT) if c is None else as_t(c)
    d = torch.zeros((p, 1), dtype=DT) if d is None else as_t(d)
    a, P = a0, P0
    F_inv_ss = None
    if kind == "steady_state":
        P_steady = solve_discrete_are(T.T, Z.T, R @ Q @ R.T, H)
        F_ss = Z @ P_steady @ Z.T + H
        F_inv_ss = torch.linalg.solve(F_ss, torch.eye(p, dtype=DT))
        P = P_steady
    step = lambda x, t: x[t] if x.ndim == 3 else x  # noqa: E731
    fs, ps, fc, pc, lls = [], [], [], [], []
    for t in range(n):
        ct, dt, Tt, Zt, Rt, Ht, Qt = (step(x, t) for x in (c, d, T, Z, R, H, Q))
        if kind == "univariate":
            a_f, a_hat, P_f, P_hat, ll = _step_univariate(data[t], a, P, ct, dt, Tt, Zt, Rt, Ht, Qt)
        else:
            y_m, Z_m, H_m, flag = _mask(data[t], Zt, Ht)
            a_f, P_f, ll = _update(kind, a, P, y_m, ct, dt, Z_m, H_m, flag, F_inv_ss, strict_reference)
            a_hat, P_hat = _predict(a_f, P_f, ct, Tt, Rt, Qt)
            if ll is None:
                ll = torch.zeros((), dtype=DT)
        if not isinstance(ll, torch.Tensor):
            ll = torch.as_tensor(ll, dtype=DT)
        fs.append(a_f), ps.append(a_hat), fc.append(P_f), pc.append(P_hat), lls.append(ll)
        a, P = a_hat, P_hat
    ll_obs = torch.stack(lls)
    return [
        torch.stack(fs),
        torch.cat([a0[None], torch.stack(ps)], 0),
        torch.stack(fc),
        torch.cat([P0[None], torch.stack(pc)], 0),
        ll_obs.sum(),
        ll_obs,
    ]


GRAD_NAMES = ("a0", "P0", "T", "Z", "R", "H", "Q", "c", "d")


def loglik_and_grads(kind, data, a0, P0, T, Z, R, H, Q, c=None, d=None, strict_reference=True, g_ll_obs=None):
    """Returns (loglik, {name: d loglik / d matrix}) via autograd; numpy in, numpy out.

    ``g_ll_obs``: optional weights w[n]; then the differentiated scalar is sum_t w_t * ll_t.
    """
    p, m = np.asarray(Z).shape[-2], np.asarray(Z).shape[-1]
    c = np.zeros((m, 1)) if c is None else c
    d = np.zeros((p, 1)) if d is None else d
    ins = {k: torch.tensor(np.asarray(v, dtype=np.float64), dtype=DT, requires_grad=True)
           for k, v in zip(GRAD_NAMES, (a0, P0, T, Z, R, H, Q, c, d))}
    out = kalman_filter(kind, data, ins["a0"], ins["P0"], ins["T"], ins["Z"], ins["R"], ins["H"], ins["Q"],
                        c=ins["c"], d=ins["d"], strict_reference=strict_reference)
    target = out[4] if g_ll_obs is None else (out[5] * torch.as_tensor(g_ll_obs, dtype=DT)).sum()
    grads = torch.autograd.grad(target, list(ins.values()), allow_unused=True)
    gd = {k: (np.zeros_like(ins[k].detach().numpy()) if g is None else g.numpy()) for k, g in zip(GRAD_NAMES, grads)}
    return float(out[4].detach()), gd


# ----------------------------------------------------------------------------
# independent check used to pin the GRADIENT oracle: autograd of the dense multivariate-normal log density
# ----------------------------------------------------------------------------
def lyapunov_dense(T, RQR):
    """Stationary P0 = T P0 T^T + R Q R^T as ONE dense linear solve, vec(P0) = (I - T (x) T)^{-1} vec(R Q R^T), written with
    differentiable torch ops only - shares neither scipy's bilinear algorithm nor the hand-written adjoint of
    ``_Lyapunov`` (the formula the reference uses, models/SARIMAX.py:106 + PyTensor's SolveDiscreteLyapunov grad)."""
    m = T.shape[0]
    A = torch.eye(m * m, dtype=DT) - torch.kron(T, T)
    return torch.linalg.solve(A, RQR.reshape(m * m, 1)).reshape(m, m)


def dense_gaussian_loglik(data, a0, P0, T, Z, R, H, Q, c=None, d=None):
    """torch twin of ``kalman_numpy.dense_gaussian_loglik``: log N(vec(y); mean, cov) of the stacked sample, no recursion
    shared with the filter, every op differentiable - ``torch.autograd`` of this scalar is an algorithm-independent
    known answer for d loglik / d (a0, P0, T, Z, R, H, Q, c, d).  Static matrices, missing entries marginalised by
    deleting rows / columns.  The covariance blocks use P0, H, Q exactly as given: for a non-symmetric P0 / H / Q only
    the symmetric part of the gradient is comparable with the filter's (DESIGN.md "gradient gauge")."""
    data = torch.as_tensor(np.asarray(data), dtype=DT)
    n, p = data.shape[0], data.shape[1]
    m = T.shape[0]
    c = torch.zeros((m, 1), dtype=DT) if c is None else c
    d = torch.zeros((p, 1), dtype=DT) if d is None else d
    RQR = R @ Q @ R.T
    means, covs = [a0], [P0]
    for _ in range(n - 1):
        means.append(T @ means[-1] + c)
        covs.append(T @ covs[-1] @ T.T + RQR)
    Tpow = [torch.eye(m, dtype=DT)]
    for _ in range(n):
        Tpow.append(T @ Tpow[-1])
    mu = torch.cat([Z @ mk + d for mk in means], dim=0).reshape(n * p)
    rows = []
    for t in range(n):
        blks = []
        for s in range(n):
            if s <= t:
                blk = Z @ (Tpow[t - s] @ covs[s]) @ Z.T  # Cov(y_t, y_s)
                if s == t:
                    blk = blk + H
            else:
                blk = (Z @ (Tpow[s - t] @ covs[t]) @ Z.T).T
            blks.append(blk)
        rows.append(torch.cat(blks, dim=1))
    S = torch.cat(rows, dim=0)
    yflat = data.reshape(n * p)
    keep = ~torch.isnan(yflat)
    idx = torch.nonzero(keep).reshape(-1)
    yv = (torch.nan_to_num(yflat) - mu)[idx]
    Sk = S[idx][:, idx]
    Sk = 0.5 * (Sk + Sk.T)
    L = torch.linalg.cholesky(Sk)
    w = torch.linalg.solve_triangular(L, yv[:, None], upper=False)[:, 0]
    return -0.5 * (idx.numel() * LOG_2PI + w @ w) - torch.log(torch.diagonal(L)).sum()


def dense_loglik_and_grads(data, a0, P0, T, Z, R, H, Q, c=None, d=None):
    """(loglik, {name: gradient}) of the dense density; same signature / return as ``loglik_and_grads``."""
    p, m = np.asarray(Z).shape[-2], np.asarray(Z).shape[-1]
    c = np.zeros((m, 1)) if c is None else c
    d = np.zeros((p, 1)) if d is None else d
    ins = {k: torch.tensor(np.asarray(v, dtype=np.float64), dtype=DT, requires_grad=True)
           for k, v in zip(GRAD_NAMES, (a0, P0, T, Z, R, H, Q, c, d))}
    ll = dense_gaussian_loglik(data, *[ins[k] for k in GRAD_NAMES])
    grads = torch.autograd.grad(ll, list(ins.values()), allow_unused=True)
    gd = {k: (np.zeros_like(ins[k].detach().numpy()) if g is None else g.numpy()) for k, g in zip(GRAD_NAMES, grads)}
    return float(ll.detach()), gd


def dare_by_riccati_iteration(T, Z, RQR, H, iters=300):
    """Stabilising solution of the filter DARE by plain fixed-point iteration of the Riccati recursion, differentiable
    torch ops only: ``torch.autograd`` through the unrolled iteration is a gradient that shares nothing with the adjoint
    formula of ``_DARE`` (utils/pytensor_scipy.py:39-60).  For contracting systems (the tests' scale of T)."""
    P = RQR.clone()
    for _ in range(iters):
        M = P @ Z.T
        P = T @ (P - M @ torch.linalg.solve(Z @ M + H, M.T)) @ T.T + RQR
        P = 0.5 * (P + P.T)
    return P
