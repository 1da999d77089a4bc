"""Seeded synthetic workloads for BASELINE.json's configs.  Host-side numpy only."""
from __future__ import annotations

import numpy as np

from .models import arma_spec, trend_seasonal_spec, varmax_spec


def simulate_arma(n, ar=(0.6,), ma=(0.3,), sigma=1.0, seed=0):
    rng = np.random.default_rng(seed)
    p, q = len(ar), len(ma)
    e = rng.normal(scale=np.sqrt(sigma), size=n + q)
    y = np.zeros(n)
    for t in range(n):
        v = e[t + q]
        for i in range(p):
            if t - 1 - i >= 0:
                v += ar[i] * y[t - 1 - i]
        for j in range(q):
            v += ma[j] * e[t + q - 1 - j]
        y[t] = v
    return y


def arma11_workload(n_draws=65536, n=1000, seed=1):
    """config[1]: BayesianARMA(1,1), stationary init; theta = [x0(2), sigma_state, rho, theta]."""
    spec = arma_spec((1, 1), stationary_initialization=True)
    y = simulate_arma(n, (0.6,), (0.3,), 1.0, seed=0)
    rng = np.random.default_rng(seed)
    theta = np.empty((n_draws, spec.n_theta))
    theta[:, 0:2] = rng.normal(size=(n_draws, 2))
    theta[:, 2] = np.exp(rng.normal(0.0, 0.3, n_draws))
    theta[:, 3] = rng.uniform(-0.95, 0.95, n_draws)
    theta[:, 4] = rng.normal(0.0, 0.5, n_draws)
    return spec, y[:, None], theta


def arma21_workload(n_draws=1 << 20, n=1000, seed=1):
    """config[4]: ARMA(2,1) sweep; theta = [x0(2), sigma_state, rho1, rho2, theta]; (rho1, rho2) from partial
    autocorrelations so every draw is stationary."""
    spec = arma_spec((2, 1), stationary_initialization=True)
    y = simulate_arma(n, (0.5, -0.2), (0.3,), 1.0, seed=0)
    rng = np.random.default_rng(seed)
    theta = np.empty((n_draws, spec.n_theta))
    theta[:, 0:2] = rng.normal(size=(n_draws, 2))
    theta[:, 2] = np.exp(rng.normal(0.0, 0.3, n_draws))
    r1, r2 = rng.uniform(-0.9, 0.9, n_draws), rng.uniform(-0.9, 0.9, n_draws)
    theta[:, 3] = r1 * (1 - r2)
    theta[:, 4] = r2
    theta[:, 5] = rng.normal(0.0, 0.5, n_draws)
    return spec, y[:, None], theta


def varmax20_workload(n_draws=262144, n=1000, k=3, missing_frac=0.1, seed=1):
    """config[2]: VARMAX(2,0), k_endog=3, measurement error, stationary init, 10% whole rows missing."""
    spec = varmax_spec(k, (2, 0), stationary_initialization=True, measurement_error=True)
    m = spec.k_states
    rng = np.random.default_rng(seed)

    def draw(nd):
        A = rng.normal(0.0, 0.2, size=(nd, k, 2 * k))
        comp = np.zeros((nd, m, m))
        comp[:, :k, :] = A
        comp[:, k:, :k] = np.eye(k)
        rad = np.abs(np.linalg.eigvals(comp)).max(axis=1)
        A *= np.minimum(1.0, 0.95 / rad)[:, None, None] ** np.array([1.0] * k + [2.0] * k)[None, None, :]
        L = np.zeros((nd, k, k))
        iu = np.tril_indices(k, -1)
        L[:, iu[0], iu[1]] = rng.normal(0.0, 0.1, size=(nd, len(iu[0])))
        L[:, np.arange(k), np.arange(k)] = np.exp(rng.normal(-1.0, 0.2, size=(nd, k)))
        Q = L @ L.transpose(0, 2, 1)
        h = np.exp(rng.normal(-2.0, 0.2, size=(nd, k)))
        x0 = rng.normal(0.0, 0.1, size=(nd, m))
        return np.concatenate([x0, A.reshape(nd, -1), Q.reshape(nd, -1), h], axis=1)

    theta = draw(n_draws)
    th0 = theta[0]
    mats = spec.matrices(th0)
    T, R, Q, H, Z = mats["T"], mats["R"], mats["Q"], mats["H"], mats["Z"]
    rs = np.random.default_rng(0)
    x = np.zeros(m)
    y = np.zeros((n, k))
    Lq = np.linalg.cholesky(Q)
    for t in range(n):
        y[t] = Z @ x + np.sqrt(np.diag(H)) * rs.normal(size=k)
        x = T @ x + R @ (Lq @ rs.normal(size=k))
    miss = np.random.default_rng(2).choice(n, int(n * missing_frac), replace=False)
    y[miss] = np.nan
    return spec, y, theta


def trend_seasonal_workload(n_draws=8192, n=2000, seed=1, period=29):
    """config[3]: trend + period-29 seasonal, k_states=30; theta = 4 variances.  (``period`` 12 -> 13 states, ...)"""
    spec = trend_seasonal_spec(period)
    rng = np.random.default_rng(seed)
    theta = np.exp(rng.normal(np.log([0.1, 0.01, 0.05, 0.5]), 0.2, size=(n_draws, 4)))
    rs = np.random.default_rng(0)
    t = np.arange(n)
    y = 0.01 * t + np.sin(2 * np.pi * t / period) + np.cumsum(rs.normal(0, 0.3, n)) + rs.normal(0, 0.7, n)
    return spec, y[:, None], theta
