"""Shared test fixtures.  make_test_inputs / nile fixture mirror reference tests/utilities/test_helpers.py:62-76,110-118."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_test_inputs(p, m, r, n, missing_data=None, H_is_zero=False, seed=0):
    data = np.arange(n * p, dtype="float").reshape(-1, p, 1)
    if missing_data is not None:
        idx = np.random.default_rng(seed).choice(n, missing_data, replace=False)
        data[idx] = np.nan
    a0 = np.zeros((m, 1))
    P0 = np.eye(m)
    Q = np.eye(r)
    H = np.zeros((p, p)) if H_is_zero else np.eye(p)
    T = np.eye(m, k=-1)
    T[0, :] = 1 / m
    R = np.eye(m)[:, :r]
    Z = np.eye(m)[:p, :]
    return data, a0, P0, T, Z, R, H, Q


def nile_data():
    return np.loadtxt(os.path.join(GOLDEN, "nile.csv"), skiprows=1).astype(float)


def nile_inputs(n_missing=0, seed=0):
    a0 = np.zeros((2, 1))
    P0 = np.eye(2) * 1e6
    Q = np.eye(2) * np.array([0.5, 0.01])
    H = np.eye(1) * 0.8
    T = np.array([[1.0, 1.0], [0.0, 1.0]])
    R = np.eye(2)
    Z = np.array([[1.0, 0.0]])
    data = nile_data()[:, None, None].copy()
    if n_missing:
        data[np.random.default_rng(seed).choice(data.shape[0], n_missing, replace=False)] = np.nan
    return data, a0, P0, T, Z, R, H, Q


def random_system(rng, m, p, r, n, n_missing=0, diag_H=False, scale_T=0.4, partial=False):
    T = rng.normal(size=(m, m)) * scale_T
    Z = rng.normal(size=(p, m))
    R = rng.normal(size=(m, r))
    A = rng.normal(size=(r, r))
    Q = A @ A.T + 0.1 * np.eye(r)
    A = rng.normal(size=(p, p))
    H = A @ A.T + 0.1 * np.eye(p)
    if diag_H:
        H = np.diag(np.diag(H))
    A = rng.normal(size=(m, m))
    P0 = A @ A.T + np.eye(m)
    a0 = rng.normal(size=(m, 1))
    y = rng.normal(size=(n, p, 1))
    if n_missing:
        y[rng.choice(n, n_missing, replace=False)] = np.nan
    if partial and p > 1:
        rows = rng.choice(n, max(1, n // 8), replace=False)
        for t in rows:
            y[t, rng.integers(0, p)] = np.nan
    return y, a0, P0, T, Z, R, H, Q


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))
