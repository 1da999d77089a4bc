"""Batched simulation helpers on the GPU (not on the logp/grad path).

Mirrors reference ``pymc_statespace/utils/simulation.py`` (numba): ``simulate_statespace`` (:29-62),
``unconditional_simulations`` (:65-79) and ``conditional_simulation`` (:17-26).  The reference draws its normals from
NumPy's global RNG inside numba, so its trajectories are not reproducible from outside; here the standard-normal draws
come from ``torch.randn`` on the device (or are passed in, which makes the result exactly checkable) and the CUDA kernels
are deterministic transforms of them.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import check, load
from .engine import _ptr, _stream_ptr


def _bs(t, base_ndim, size):
    return size if t.ndim == base_ndim + 1 else 0


def simulate_statespace(T, Z, R, H, Q, n_steps: int, x0=None, n_simulations: int = 1, z_state=None, z_obs=None,
                        generator: Optional[torch.Generator] = None):
    """T:[B,m,m]|[m,m] ... float64 CUDA tensors.  Returns (states[B*S,n,m], obs[B*S,n,p]) for S = n_simulations
    trajectories per draw (``unconditional_simulations`` ordering: draw-major)."""
    lib = load()
    dev = T.device
    m, p, r = T.shape[-1], Z.shape[-2], R.shape[-1]
    B = max([t.shape[0] for t, nd in ((T, 2), (Z, 2), (R, 2), (H, 2), (Q, 2)) if t.ndim == nd + 1] + [1])
    S = B * n_simulations
    T, Z, R, H, Q = (t.contiguous() for t in (T, Z, R, H, Q))
    if z_state is None:
        z_state = torch.randn((S, n_steps, r), dtype=torch.float64, device=dev, generator=generator)
    if z_obs is None:
        z_obs = torch.randn((S, n_steps, p), dtype=torch.float64, device=dev, generator=generator)
    z_state, z_obs = z_state.contiguous(), z_obs.contiguous()
    if tuple(z_state.shape) != (S, n_steps, r) or tuple(z_obs.shape) != (S, n_steps, p):
        raise ValueError("z_state / z_obs must have shapes [B*S, n, k_posdef] / [B*S, n, k_endog]")
    x0c = None if x0 is None else x0.reshape(-1, m).contiguous()
    states = torch.empty((S, n_steps, m), dtype=torch.float64, device=dev)
    obs = torch.empty((S, n_steps, p), dtype=torch.float64, device=dev)
    info = torch.zeros(S, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.kfb_simulate(B, n_simulations, n_steps, m, p, r, _ptr(T), _bs(T, 2, m * m), _ptr(Z), _bs(Z, 2, p * m),
                               _ptr(R), _bs(R, 2, m * r), _ptr(H), _bs(H, 2, p * p), _ptr(Q), _bs(Q, 2, r * r), _ptr(x0c),
                               0 if x0c is None or x0c.shape[0] == 1 else m, _ptr(z_state), _ptr(z_obs), _ptr(states),
                               _ptr(obs), _ptr(info), _stream_ptr(dev)), "kfb_simulate")
    if int(info.sum()) != 0:
        raise RuntimeError("simulate_statespace: Q or H is not positive definite for some draw")
    return states, obs


def conditional_simulation(mus, covs, n_simulations: int = 100, z=None, jitter=None,
                           generator: Optional[torch.Generator] = None):
    """mus[U,n,k], covs[U,n,k,k] (filtered / predicted / smoothed moments) -> simulations[U*S,n,k]:
    each time block is an independent draw mu_t + chol(cov_t + jitter I) z (reference :8-26; the reference's jitter is
    uniform(1e-12, 1e-8) per trajectory)."""
    lib = load()
    dev = mus.device
    U, n, k = mus.shape
    S = U * n_simulations
    if z is None:
        z = torch.randn((S, n, k), dtype=torch.float64, device=dev, generator=generator)
    if jitter is None:
        jitter = 1e-12 + (1e-8 - 1e-12) * torch.rand(S, dtype=torch.float64, device=dev, generator=generator)
    mus, covs, z, jitter = mus.contiguous(), covs.contiguous(), z.contiguous(), jitter.contiguous()
    out = torch.empty((S, n, k), dtype=torch.float64, device=dev)
    info = torch.zeros(S, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.kfb_mvn_draws(U, n_simulations, n, k, _ptr(mus), _ptr(covs), _ptr(z), _ptr(jitter), _ptr(out), _ptr(info),
                                _stream_ptr(dev)), "kfb_mvn_draws")
    if int(info.sum()) != 0:
        raise RuntimeError("conditional_simulation: a covariance block is not positive definite")
    return out
