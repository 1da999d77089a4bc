"""Literal NumPy/SciPy restatement of the reference Kalman filters (ORACLE - test infrastructure).

Follows ``/root/reference/pymc_statespace/filters/kalman_filter.py`` operation by
operation (same association order of the matrix products, same LAPACK-backed
SciPy routines PyTensor dispatches to), one Python-level step per time step -
the closest CPU analogue of the ``pytensor.scan`` the reference runs.

Shapes at the seam (``kalman_filter.py:126-128``, ``tests/utilities/test_helpers.py:27``):
``data[n,p,1]`` (NaN = missing), ``a0[m,1]``, ``P0[m,m]``, ``T[m,m]``, ``Z[p,m]``,
``R[m,r]``, ``H[p,p]``, ``Q[r,r]``, optional ``c[m,1]``, ``d[p,1]``; any of
c,d,T,Z,R,H,Q may be 3-D time-first (``filters/utilities.py:9-14``).

Returns the reference's 6-list (``kalman_filter.py:184-191``):
``filtered_states[n,m,1], predicted_states[n+1,m,1], filtered_covs[n,m,m],
predicted_covs[n+1,m,m], loglike (scalar), ll_obs[n]``.

``strict_reference=True`` reproduces the quirks listed in SURVEY.md A.2
(Q1 single log(2 pi) in standard/steady_state, Q4 diag-only second triangular
solve in cholesky, Q5 sign of d in single, Q6 no d in steady_state).
``strict_reference=False`` gives the mathematically intended filter (used for
cross-filter identities in the tests).
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.linalg

LOG_2PI = float(np.log(2.0 * np.pi))  # MVN_CONST, kalman_filter.py:16

FILTER_KINDS = ("standard", "univariate", "steady_state", "single", "cholesky")


# ----------------------------------------------------------------------------
# helpers shared by all filters
# ----------------------------------------------------------------------------
def _as_step(mat, t):
    """Static (2-D) or time-first 3-D matrix -> value at step t (filters/utilities.py:1-20)."""
    return mat[t] if mat.ndim == 3 else mat


def handle_missing_values(y, Z, H):
    """kalman_filter.py:196-213 - rows of Z and ROWS of H are zeroed, y NaN -> 0."""
    nan_mask = np.isnan(y)
    all_nan_flag = float(np.all(nan_mask))
    W = np.eye(y.shape[0])
    idx = nan_mask.ravel()
    W[idx, idx] = 0.0
    Z_masked = W.dot(Z)
    H_masked = W.dot(H)
    y_masked = y.copy()
    y_masked[nan_mask] = 0.0
    return y_masked, Z_masked, H_masked, all_nan_flag


def predict(a, P, c, T, R, Q):
    """kalman_filter.py:216-223."""
    a_hat = T.dot(a) + c
    P_hat = T.dot(P).dot(T.T) + R.dot(Q).dot(R.T)
    P_hat = 0.5 * (P_hat + P_hat.T)
    return a_hat, P_hat


def _solve_pos(F, B):
    # pt.linalg.solve(..., assume_a="pos") -> scipy.linalg.solve -> LAPACK posv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", scipy.linalg.LinAlgWarning)
        return scipy.linalg.solve(F, B, assume_a="pos", check_finite=False)


def _log_det(F):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.linalg.det(F))


# ----------------------------------------------------------------------------
# the five update rules
# ----------------------------------------------------------------------------
def update_standard(a, P, y, c, d, Z, H, all_nan_flag, strict_reference=True):
    """StandardFilter.update, kalman_filter.py:255-284."""
    m, p = P.shape[0], Z.shape[0]
    eye_endog, eye_states = np.eye(p), np.eye(m)
    v = y - Z.dot(a) - d
    PZT = P.dot(Z.T)
    F = Z.dot(PZT) + H
    F_inv = _solve_pos(F + eye_endog * all_nan_flag, eye_endog)
    K = PZT.dot(F_inv)
    I_KZ = eye_states - K.dot(Z)
    a_filtered = a + K.dot(v)
    P_filtered = I_KZ.dot(P).dot(I_KZ.T) + K.dot(H).dot(K.T)
    inner_term = v.T.dot(F_inv).dot(v)
    n_const = 1.0 if strict_reference else float(p)  # Q1: log(2 pi) counted once
    if all_nan_flag:
        ll = 0.0
    else:
        ll = float((-0.5 * (n_const * LOG_2PI + _log_det(F) + inner_term)).ravel()[0])
    return a_filtered, P_filtered, ll


def update_cholesky(a, P, y, c, d, Z, H, all_nan_flag, strict_reference=True):
    """CholeskyFilter.update, kalman_filter.py:287-318.

    strict: the second ``SolveTriangular(lower=True)`` is handed ``F_chol.T`` (upper
    triangular), so LAPACK trtrs reads only its lower triangle = diag(L).
    """
    m, p = P.shape[0], Z.shape[0]
    eye_endog, eye_states = np.eye(p), np.eye(m)
    v = y - Z.dot(a) - d
    PZT = P.dot(Z.T)
    F = Z.dot(PZT) + H + eye_endog * all_nan_flag
    F_chol = scipy.linalg.cholesky(F, lower=True, check_finite=False)
    second_lower = bool(strict_reference)

    def solve_lower(A, b):
        return scipy.linalg.solve_triangular(A, b, lower=True, check_finite=False)

    def solve_second(A, b):
        return scipy.linalg.solve_triangular(A, b, lower=second_lower, check_finite=False)

    K = solve_second(F_chol.T, solve_lower(F_chol, PZT.T)).T * (1.0 - all_nan_flag)
    I_KZ = eye_states - K.dot(Z)
    a_filtered = a + K.dot(v)
    P_filtered = I_KZ.dot(P).dot(I_KZ.T) + K.dot(H).dot(K.T)
    inner_term = solve_second(F_chol.T, solve_lower(F_chol, v))
    n = y.shape[0]
    if all_nan_flag:
        ll = 0.0
    else:
        ll = float(
            (-0.5 * (n * LOG_2PI + (v.T @ inner_term).ravel()) - np.log(np.diag(F_chol)).sum()).ravel()[0]
        )
    return a_filtered, P_filtered, ll


def update_single(a, P, y, c, d, Z, H, all_nan_flag, strict_reference=True):
    """SingleTimeseriesFilter.update, kalman_filter.py:333-351 (Q5: v = y - (Za - d))."""
    m = P.shape[0]
    eye_states = np.eye(m)
    if strict_reference:
        y_hat = Z.dot(a).ravel() - d
    else:
        y_hat = Z.dot(a).ravel() + d
    v = y - y_hat
    PZT = P.dot(Z.T)
    F = (Z.dot(PZT) + H).ravel() + all_nan_flag
    K = PZT / F
    I_KZ = eye_states - K.dot(Z)
    a_filtered = a + (K * v)
    P_filtered = I_KZ.dot(P).dot(I_KZ.T) + K.dot(H).dot(K.T)
    if all_nan_flag:
        ll = 0.0
    else:
        ll = float((-0.5 * (LOG_2PI + np.log(F) + v**2 / F)).ravel()[0])
    return a_filtered, P_filtered, ll


def update_steady_state(a, P, c, d, F_inv, y, Z, H, all_nan_flag, strict_reference=True):
    """SteadyStateFilter.update, kalman_filter.py:399-419 (Q6: d ignored)."""
    m, p = P.shape[0], Z.shape[0]
    eye_states = np.eye(m)
    v = y - Z.dot(a)
    if not strict_reference:
        v = v - d
    PZT = P.dot(Z.T)
    F = Z.dot(PZT) + H
    K = PZT.dot(F_inv)
    I_KZ = eye_states - K.dot(Z)
    a_filtered = a + K.dot(v)
    P_filtered = I_KZ.dot(P).dot(I_KZ.T) + K.dot(H).dot(K.T)
    inner_term = v.T.dot(F_inv).dot(v)
    n_const = 1.0 if strict_reference else float(p)
    if all_nan_flag:
        ll = 0.0
    else:
        ll = float((-0.5 * (n_const * LOG_2PI + _log_det(F) + inner_term)).ravel()[0])
    return a_filtered, P_filtered, ll


def univariate_inner_step(y, Z_row, d_row, sigma_H, nan_flag, a, P):
    """UnivariateFilter._univariate_inner_filter_step, kalman_filter.py:460-480."""
    Z_row = Z_row[None, :]
    v = y - Z_row.dot(a) - d_row
    PZT = P.dot(Z_row.T)
    F = Z_row.dot(PZT) + sigma_H
    F_zero_flag = np.logical_or(F == 0, nan_flag)
    F = F + 1e-8 * F_zero_flag
    keep = 1.0 - F_zero_flag
    K = PZT / F * keep
    a_filtered = a + K * v * keep
    P_filtered = P - np.outer(K, K) * F * keep
    ll_inner = (np.log(F) + v**2 / F) * keep
    return a_filtered, P_filtered, ll_inner


def step_univariate(y, a, P, c, d, T, Z, R, H, Q):
    """UnivariateFilter.kalman_step, kalman_filter.py:482-505."""
    y = y[:, None]  # [p,1,1]
    nan_mask = np.isnan(y).ravel()
    W = np.eye(y.shape[0])
    W[nan_mask, nan_mask] = 0.0
    Z_masked = W.dot(Z)
    H_masked = W.dot(H)
    y_masked = y.copy()
    y_masked[nan_mask] = 0.0
    sigma = np.diag(H_masked)
    ll_inner = []
    for i in range(y.shape[0]):
        a, P, lli = univariate_inner_step(y_masked[i], Z_masked[i], d[i], sigma[i], nan_mask[i], a, P)
        ll_inner.append(lli)
    ll_inner = np.stack(ll_inner)
    a_filtered, P_filtered = a, P
    a_hat, P_hat = predict(a_filtered, P_filtered, c, T, R, Q)
    ll = float(-0.5 * ((ll_inner != 0).sum() * LOG_2PI + ll_inner.sum()))
    return a_filtered, a_hat, P_filtered, P_hat, ll


_UPDATES = {"standard": update_standard, "cholesky": update_cholesky, "single": update_single}


# ----------------------------------------------------------------------------
# the scan
# ----------------------------------------------------------------------------
def kalman_filter(kind, data, a0, P0, T, Z, R, H, Q, c=None, d=None, strict_reference=True):
    """BaseFilter.build_graph + _postprocess_scan_results (kalman_filter.py:126-193)."""
    kind = kind.lower()
    if kind not in FILTER_KINDS:
        raise NotImplementedError("The following are valid filter types: " + ", ".join(FILTER_KINDS))
    data = np.asarray(data, dtype=np.float64)
    a0, P0, T, Z, R, H, Q = (np.asarray(x, dtype=np.float64) for x in (a0, P0, T, Z, R, H, Q))
    n = data.shape[0]
    k_endog, k_states = Z.shape[-2], Z.shape[-1]
    if c is None:
        c = np.zeros((k_states, 1))  # initialize_intercepts, kalman_filter.py:37-51
    if d is None:
        d = np.zeros((k_endog, 1))
    c, d = np.asarray(c, dtype=np.float64), np.asarray(d, dtype=np.float64)
    if kind == "single" and data.shape[1] != 1:
        # assert_data_is_1d, kalman_filter.py:19,329
        raise AssertionError("UnivariateTimeSeries filter requires data be at most 1-dimensional")
    for name, mat in zip(("c", "d", "T", "Z", "R", "H", "Q"), (c, d, T, Z, R, H, Q)):
        if mat.ndim == 3 and mat.shape[0] != n:
            raise AssertionError(
                "The first dimension of a time varying matrix (the time dimension) must be "
                "equal to the first dimension of the data (the time dimension)."
            )
        if mat.ndim not in (2, 3):
            raise ValueError(f"Matrix {name} has {mat.ndim}, it should either 2 (static) or 3 (time varying).")

    a, P = a0, P0
    F_inv_ss = None
    if kind == "steady_state":
        # SteadyStateFilter.build_graph, kalman_filter.py:384-391 (static matrices only)
        P_steady = scipy.linalg.solve_discrete_are(T.T, Z.T, R.dot(Q).dot(R.T), H)
        F_ss = Z.dot(P_steady).dot(Z.T) + H
        F_inv_ss = _solve_pos(F_ss, np.eye(F_ss.shape[0]))
        P = P_steady

    fs, ps, fc, pc, lls = [], [], [], [], []
    for t in range(n):
        y = data[t]
        ct, dt, Tt, Zt, Rt, Ht, Qt = (_as_step(x, t) for x in (c, d, T, Z, R, H, Q))
        if kind == "univariate":
            a_f, a_hat, P_f, P_hat, ll = step_univariate(y, a, P, ct, dt, Tt, Zt, Rt, Ht, Qt)
        else:
            y_m, Z_m, H_m, flag = handle_missing_values(y, Zt, Ht)
            if kind == "steady_state":
                a_f, P_f, ll = update_steady_state(a, P, ct, dt, F_inv_ss, y_m, Z_m, H_m, flag, strict_reference)
            else:
                a_f, P_f, ll = _UPDATES[kind](a, P, y_m, ct, dt, Z_m, H_m, flag, strict_reference)
            a_hat, P_hat = predict(a_f, P_f, ct, Tt, Rt, Qt)
        fs.append(a_f), ps.append(a_hat), fc.append(P_f), pc.append(P_hat), lls.append(ll)
        a, P = a_hat, P_hat

    filtered_states = np.stack(fs)
    predicted_states = np.concatenate([a0[None], np.stack(ps)], axis=0)
    filtered_covs = np.stack(fc)
    predicted_covs = np.concatenate([P0[None], np.stack(pc)], axis=0)
    ll_obs = np.asarray(lls, dtype=np.float64)
    return [filtered_states, predicted_states, filtered_covs, predicted_covs, float(ll_obs.sum()), ll_obs]


# ----------------------------------------------------------------------------
# independent check used to pin the oracle: dense multivariate-normal log density
# ----------------------------------------------------------------------------
def dense_gaussian_loglik(data, a0, P0, T, Z, R, H, Q, c=None, d=None, mp_digits=None):
    """log N(vec(y); mean, cov) of the whole sample from the state-space moments (no recursion
    shared with the filter).  Static matrices.  Missing observations (NaN entries; whole rows or single entries) are
    marginalised exactly by deleting their rows / columns from the stacked mean and covariance.
    O((n p)^3) - small cases only (the Nile fixture, n = 100, is a 100 x 100 Cholesky).
    ``mp_digits``: factorise with mpmath at that many decimal digits - the stacked covariance of a diffuse start
    (P0 = 1e6 I) has a condition number ~1e12 and a float64 Cholesky of it is only good to ~1e-6."""
    data = np.asarray(data, dtype=np.float64)
    n, p = data.shape[0], data.shape[1]
    m = T.shape[0]
    c = np.zeros((m, 1)) if c is None else c
    d = np.zeros((p, 1)) if d is None else d
    if mp_digits:
        return _dense_gaussian_loglik_mp(data, a0, P0, T, Z, R, H, Q, c, d, int(mp_digits))
    RQR = R @ Q @ R.T
    means, covs = [a0], [P0]
    for _ in range(n - 1):
        means.append(T @ means[-1] + c)
        covs.append(T @ covs[-1] @ T.T + RQR)
    Tpow = [np.eye(m)]
    for _ in range(n):
        Tpow.append(T @ Tpow[-1])
    mu = np.concatenate([Z @ mk + d for mk in means], axis=0).ravel()
    S = np.zeros((n * p, n * p))
    for s in range(n):
        for t in range(s, n):
            blk = Z @ (Tpow[t - s] @ covs[s]) @ Z.T  # Cov(y_t, y_s)
            if s == t:
                blk = blk + H
            S[t * p : (t + 1) * p, s * p : (s + 1) * p] = blk
            S[s * p : (s + 1) * p, t * p : (t + 1) * p] = blk.T
    yflat = data.reshape(n * p)
    keep = ~np.isnan(yflat)
    yv = (yflat - mu)[keep]
    L = np.linalg.cholesky(S[np.ix_(keep, keep)])
    w = scipy.linalg.solve_triangular(L, yv, lower=True)
    return float(-0.5 * (int(keep.sum()) * LOG_2PI + w @ w) - np.log(np.diag(L)).sum())


def _dense_gaussian_loglik_mp(data, a0, P0, T, Z, R, H, Q, c, d, digits):
    """dense_gaussian_loglik with every operation (moments, covariance blocks, Cholesky) in mpmath."""
    import mpmath as mp

    with mp.workdps(digits):
        M = lambda x: mp.matrix(np.asarray(x, dtype=np.float64).tolist())  # noqa: E731
        n, p = data.shape[0], data.shape[1]
        m = T.shape[0]
        a0, P0, T, Z, R, H, Q, c, d = (M(np.asarray(x).reshape(np.asarray(x).shape[0], -1)) for x in (a0, P0, T, Z, R, H, Q, c, d))
        RQR = R * Q * R.T
        means, covs = [a0], [P0]
        for _ in range(n - 1):
            means.append(T * means[-1] + c)
            covs.append(T * covs[-1] * T.T + RQR)
        Tpow = [mp.eye(m)]
        for _ in range(n):
            Tpow.append(T * Tpow[-1])
        yflat = np.asarray(data, dtype=np.float64).reshape(n * p)
        keep = [i for i in range(n * p) if not np.isnan(yflat[i])]
        pos = {i: k for k, i in enumerate(keep)}
        S = mp.zeros(len(keep), len(keep))
        for s_ in range(n):
            for t in range(s_, n):
                blk = Z * (Tpow[t - s_] * covs[s_]) * Z.T
                if s_ == t:
                    blk = blk + H
                for i in range(p):
                    for j in range(p):
                        gi, gj = t * p + i, s_ * p + j
                        if gi in pos and gj in pos:
                            S[pos[gi], pos[gj]] = blk[i, j]
                            S[pos[gj], pos[gi]] = blk[i, j]
        mu = [(Z * means[t] + d)[i] for t in range(n) for i in range(p)]
        yv = mp.matrix([mp.mpf(float(yflat[i])) - mu[i] for i in keep])
        L = mp.cholesky(S)
        w = mp.lu_solve(L, yv)
        quad = sum(x * x for x in w)
        logdet = sum(mp.log(L[i, i]) for i in range(L.rows))
        return float(-(len(keep) * mp.log(2 * mp.pi) + quad) / 2 - logdet)


def dense_gaussian_loglik_time_varying(data, a0, P0, T, Z, R, H, Q, c=None, d=None):
    """``dense_gaussian_loglik`` for time-varying matrices (any of T, Z, R, H, Q, c, d may be 3-D, time first, as
    filters/utilities.py:9-14 slices them): x_{t+1} = T_t x_t + c_t + R_t eta_t (Q_t), y_t = Z_t x_t + d_t + eps_t (H_t).
    Dense algebra only; whole-row / single-entry missing values are deleted."""
    data = np.asarray(data, dtype=np.float64)
    n, p = data.shape[0], data.shape[1]
    at = lambda M, t: None if M is None else (M[t] if np.ndim(M) == 3 else M)  # noqa: E731
    m = at(T, 0).shape[0]
    zc, zd = np.zeros((m, 1)), np.zeros((p, 1))
    ct = lambda t: zc if c is None else at(c, t)  # noqa: E731
    dt = lambda t: zd if d is None else at(d, t)  # noqa: E731
    means, covs = [a0], [P0]
    for t in range(n - 1):
        Tt, Rt = at(T, t), at(R, t)
        means.append(Tt @ means[-1] + ct(t))
        covs.append(Tt @ covs[-1] @ Tt.T + Rt @ at(Q, t) @ Rt.T)
    mu = np.concatenate([at(Z, t) @ means[t] + dt(t) for t in range(n)], axis=0).ravel()
    S = np.zeros((n * p, n * p))
    for s in range(n):
        Phi = np.eye(m)  # T_{t-1} ... T_s
        for t in range(s, n):
            blk = at(Z, t) @ (Phi @ covs[s]) @ at(Z, s).T  # Cov(y_t, y_s)
            if s == t:
                blk = blk + at(H, t)
            S[t * p : (t + 1) * p, s * p : (s + 1) * p] = blk
            S[s * p : (s + 1) * p, t * p : (t + 1) * p] = blk.T
            Phi = at(T, t) @ Phi
    yflat = data.reshape(n * p)
    keep = ~np.isnan(yflat)
    yv = (yflat - mu)[keep]
    L = np.linalg.cholesky(S[np.ix_(keep, keep)])
    w = scipy.linalg.solve_triangular(L, yv, lower=True)
    return float(-0.5 * (int(keep.sum()) * LOG_2PI + w @ w) - np.log(np.diag(L)).sum())


def dense_gaussian_state_moments(data, a0, P0, T, Z, R, H, Q, c=None, d=None):
    """Conditional moments of the states from the joint Gaussian of (x_0..x_n, y_0..y_{n-1}), by dense linear algebra
    only (no recursion shared with the filter or the smoother): returns a function ``cond(t, s)`` giving
    (E[x_t | y_0..y_s], Cov[x_t | y_0..y_s]); s = t: filtered, t = s + 1: predicted, s = n - 1: smoothed.  x_0 ~ N(a0, P0) is
    the predicted state of the first step (kalman_filter.py:235-252).  Static matrices; missing entries are dropped from
    the conditioning set.  Small cases only."""
    data = np.asarray(data, dtype=np.float64)
    n, p = data.shape[0], data.shape[1]
    m = T.shape[0]
    c = np.zeros((m, 1)) if c is None else c
    d = np.zeros((p, 1)) if d is None else d
    RQR = R @ Q @ R.T
    means, covs = [a0], [P0]
    for _ in range(n):
        means.append(T @ means[-1] + c)
        covs.append(T @ covs[-1] @ T.T + RQR)
    Tpow = [np.eye(m)]
    for _ in range(n + 1):
        Tpow.append(T @ Tpow[-1])

    def cov_xx(t, s):  # Cov(x_t, x_s)
        return Tpow[t - s] @ covs[s] if t >= s else (Tpow[s - t] @ covs[t]).T

    def cond(t, s):
        idx = [(k, i) for k in range(s + 1) for i in range(p) if not np.isnan(data[k, i, 0])]
        if not idx:
            return means[t], covs[t]
        Zr = lambda i: Z[i : i + 1, :]  # noqa: E731
        Syy = np.array([[(Zr(i) @ cov_xx(k, l) @ Zr(j).T)[0, 0] + (H[i, j] if k == l else 0.0) for (l, j) in idx]
                        for (k, i) in idx])
        Sxy = np.concatenate([cov_xx(t, k) @ Zr(i).T for (k, i) in idx], axis=1)
        resid = np.array([[data[k, i, 0] - (Zr(i) @ means[k])[0, 0] - d[i, 0]] for (k, i) in idx])
        W = np.linalg.solve(Syy, Sxy.T).T
        return means[t] + W @ resid, covs[t] - W @ Sxy.T

    return cond


# ----------------------------------------------------------------------------
# RTS smoother (SURVEY.md section 8(f) row f2 - not on the logp/grad path)
# ----------------------------------------------------------------------------
def kalman_smoother(T, R, Q, filtered_states, filtered_covariances):
    """KalmanSmoother.build_graph / smoother_step, reference filters/kalman_smoother.py:56-104 (static T, R, Q).

    Backwards scan from the last filtered moment; the gain uses ``pinv(P_hat)`` (:92) and ``predict`` here has no
    intercept and no symmetrisation (:99-104).  Returns (smoothed_states[n,m,1], smoothed_covariances[n,m,m])."""
    T, R, Q = (np.asarray(x, dtype=np.float64) for x in (T, R, Q))
    fs = np.asarray(filtered_states, dtype=np.float64)
    fc = np.asarray(filtered_covariances, dtype=np.float64)
    n = fs.shape[0]
    a_smooth, P_smooth = fs[-1], fc[-1]
    out_a, out_P = [a_smooth], [P_smooth]
    for t in range(n - 2, -1, -1):
        a, P = fs[t], fc[t]
        a_hat = T.dot(a)
        P_hat = T.dot(P).dot(T.T) + R.dot(Q).dot(R.T)
        smoother_gain = np.linalg.pinv(P_hat).dot(T).dot(P).T
        a_smooth = a + smoother_gain @ (a_smooth - a_hat)
        P_smooth = P + smoother_gain.dot(P_smooth - P_hat).dot(smoother_gain.T)
        out_a.append(a_smooth)
        out_P.append(P_smooth)
    return np.stack(out_a[::-1]), np.stack(out_P[::-1])
