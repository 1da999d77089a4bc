"""theta -> matrices of the reference models, restated with torch ops (ORACLE - test infrastructure).

Follows ``update()`` of reference ``pymc_statespace/models/SARIMAX.py:59-107``, ``models/VARMAX.py:95-150``
and ``models/local_level.py:28-49`` statement by statement (cursor arithmetic included), so the product's
declarative index maps (``pymc_statespace_b200/models.py``) are checked against an independent restatement.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kalman_torch as kt

DT = torch.float64


def _z(*shape):
    return torch.zeros(shape, dtype=DT)


def arma_matrices(theta, order, stationary_initialization=True):
    p, q = order
    k_states = max(p, q + 1)
    Z = torch.tensor(np.r_[[1.0], np.zeros(k_states - 1)][None], dtype=DT)
    T = torch.tensor(np.eye(k_states, k=1), dtype=DT)
    R = torch.tensor(np.r_[[[1.0]], np.zeros(k_states - 1)[:, None]], dtype=DT)
    a0 = _z(k_states, 1)
    P0 = torch.eye(k_states, dtype=DT)
    H = _z(1, 1)
    Q = _z(1, 1)
    cursor = 0
    a0 = a0.clone()
    a0[:, 0] = theta[cursor:cursor + k_states]
    cursor += k_states
    if not stationary_initialization:
        P0 = theta[cursor:cursor + k_states**2].reshape(k_states, k_states)
        cursor += k_states**2
    Q = Q.clone()
    Q[0, 0] = theta[cursor]
    cursor += 1
    T = T.clone()
    T[np.arange(p), np.zeros(p, dtype=int)] = theta[cursor:cursor + p]
    cursor += p
    R = R.clone()
    R[np.arange(1, q + 1), np.zeros(q, dtype=int)] = theta[cursor:cursor + q]
    cursor += q
    if stationary_initialization:
        P0 = kt.solve_discrete_lyapunov(T, R @ Q @ R.T)
    return a0, P0, T, Z, R, H, Q


def varmax_matrices(theta, k_obs, order, stationary_initialization=True, measurement_error=True):
    p, q = order
    k_order = max(p, 1) + q
    k_states = k_obs * k_order
    k_posdef = k_obs
    Z = _z(k_obs, k_states)
    Z[np.arange(k_obs), np.arange(k_obs)] = 1
    T = _z(k_states, k_states)
    if p > 1:
        T[k_obs:k_obs * p, 0:k_obs * (p - 1)] = torch.eye(k_obs * (p - 1), dtype=DT)
    if q > 1:
        T[-k_obs * (q - 1):, -k_obs * q:-k_obs] = torch.eye(k_obs * (q - 1), dtype=DT)
    R = _z(k_states, k_obs)
    R[0:k_obs, :] = torch.eye(k_obs, dtype=DT)
    if q > 0:
        end = -k_obs * (q - 1) if q > 1 else None
        R[slice(k_obs * -q, end), :] = torch.eye(k_obs, dtype=DT)
    a0, P0, H, Q = _z(k_states, 1), _z(k_states, k_states), _z(k_obs, k_obs), _z(k_posdef, k_posdef)
    cursor = 0
    a0[:, 0] = theta[cursor:cursor + k_states]
    cursor += k_states
    if not stationary_initialization:
        P0 = theta[cursor:cursor + k_states**2].reshape(k_states, k_states)
        cursor += k_states**2
    if p > 0:
        cnt = k_obs**2 * p
        T[0:k_obs, 0:k_obs * p] = theta[cursor:cursor + cnt].reshape(k_obs, k_obs * p)
        cursor += cnt
    if q > 0:
        cnt = k_obs**2 * q
        T[0:k_obs, k_obs * max(1, p):] = theta[cursor:cursor + cnt].reshape(k_obs, k_obs * q)
        cursor += cnt
    Q = theta[cursor:cursor + k_posdef**2].reshape(k_posdef, k_posdef)
    cursor += k_posdef**2
    if measurement_error:
        H[np.arange(k_obs), np.arange(k_obs)] = theta[cursor:cursor + k_obs]
        cursor += k_obs
    if stationary_initialization:
        P0 = kt.solve_discrete_lyapunov(T, R @ Q @ R.T)
    return a0, P0, T, Z, R, H, Q


def local_level_matrices(theta):
    Z = torch.tensor([[1.0, 0.0]], dtype=DT)
    T = torch.tensor([[1.0, 1.0], [0.0, 1.0]], dtype=DT)
    R = torch.eye(2, dtype=DT)
    a0 = _z(2, 1)
    a0[:, 0] = theta[:2]
    P0 = theta[2:6].reshape(2, 2)
    H = _z(1, 1)
    H[0, 0] = theta[6]
    Q = _z(2, 2)
    Q[np.arange(2), np.arange(2)] = theta[7:]
    return a0, P0, T, Z, R, H, Q


def logp_and_grad_theta(matrices_fn, theta, data, kind="standard", strict_reference=True):
    """(logp, dlogp/dtheta) for ONE theta through the torch twin.  data: [n,p,1] numpy."""
    th = torch.tensor(np.asarray(theta, dtype=np.float64), dtype=DT, requires_grad=True)
    a0, P0, T, Z, R, H, Q = matrices_fn(th)
    out = kt.kalman_filter(kind, torch.as_tensor(np.asarray(data), dtype=DT), a0, P0, T, Z, R, H, Q,
                           strict_reference=strict_reference)
    (g,) = torch.autograd.grad(out[4], [th])
    return float(out[4].detach()), g.numpy()
