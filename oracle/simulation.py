"""Restatement of reference pymc_statespace/utils/simulation.py with the normal draws passed in (ORACLE)."""
import numpy as np


def simulate_statespace(T, Z, R, H, Q, n_steps, z_state, z_obs, x0=None):
    """simulation.py:29-62; z_state[n,r], z_obs[n,p] replace the two np.random.randn calls."""
    n_obs, n_states = Z.shape
    k_obs_noise = H.shape[0] * (1 - int(np.all(H == 0)))
    state_innovations = z_state @ np.linalg.cholesky(Q)
    if k_obs_noise != 0:
        obs_innovations = z_obs @ np.linalg.cholesky(H)
    simulated_states = np.zeros((n_steps, n_states))
    simulated_obs = np.zeros((n_steps, n_obs))
    if x0 is not None:
        simulated_states[0] = x0
        simulated_obs[0] = Z @ x0
    for t in range(1, n_steps):
        simulated_states[t] = T @ simulated_states[t - 1] + R @ state_innovations[t]
        simulated_obs[t] = Z @ simulated_states[t - 1] + (obs_innovations[t] if k_obs_noise != 0 else 0.0)
    return simulated_states, simulated_obs


def mvn_draws_blockdiag(mu, covs, z, jitter):
    """numba_mvn_draws on numba_block_diagonal(covs) (:8-26): mu[n,k], covs[n,k,k], z[n,k]."""
    n, k = mu.shape
    big = np.zeros((n * k, n * k))
    for t in range(n):
        big[t * k:(t + 1) * k, t * k:(t + 1) * k] = covs[t]
    L = np.linalg.cholesky(big + np.eye(n * k) * jitter)
    return (mu.reshape(-1) + L @ z.reshape(-1)).reshape(n, k)
