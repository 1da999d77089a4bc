"""Row f4: batched simulation kernels vs the restatement of reference utils/simulation.py fed the same normal draws."""
import numpy as np
import pytest
import torch

from oracle import kalman_numpy as kn
from oracle import simulation as osim
from tests.helpers import random_system, rel_err

pytestmark = pytest.mark.gpu


def _dev(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("h_zero", [False, True])
def test_simulate_statespace_matches_restatement(h_zero):
    from pymc_statespace_b200.simulation import simulate_statespace

    rng = np.random.default_rng(3)
    B, S, n, m, p, r = 5, 3, 40, 4, 2, 2
    systems = [random_system(rng, m, p, r, n) for _ in range(B)]
    stack = lambda i: np.stack([s[i] for s in systems])  # noqa: E731
    T, Z, R, H, Q = stack(3), stack(4), stack(5), stack(6), stack(7)
    if h_zero:
        H = np.zeros_like(H)
    x0 = rng.normal(size=(B, m))
    zs, zo = rng.normal(size=(B * S, n, r)), rng.normal(size=(B * S, n, p))
    states, obs = simulate_statespace(_dev(T), _dev(Z), _dev(R), _dev(H), _dev(Q), n, x0=_dev(x0), n_simulations=S,
                                      z_state=_dev(zs), z_obs=_dev(zo))
    states, obs = states.cpu().numpy(), obs.cpu().numpy()
    for s in (0, 4, 14):
        b = s // S
        rs, ro = osim.simulate_statespace(T[b], Z[b], R[b], H[b], Q[b], n, zs[s], zo[s], x0=x0[b])
        assert rel_err(states[s], rs) < 1e-12 and rel_err(obs[s], ro) < 1e-12
    # shared matrices, no x0, internally generated noise: shapes + second moments
    st, ob = simulate_statespace(_dev(T[0] * 0.5), _dev(Z[0]), _dev(R[0]), _dev(H[0]), _dev(Q[0]), 60, n_simulations=4000,
                                 generator=torch.Generator(device="cuda").manual_seed(1))
    assert st.shape == (4000, 60, m) and ob.shape == (4000, 60, p)
    assert float(st[:, 0].abs().max()) == 0.0
    import scipy.linalg

    L = np.linalg.cholesky(Q[0])
    cov_innov = L.T @ L  # the reference multiplies the row vector by the lower factor (:37-38)
    P_inf = scipy.linalg.solve_discrete_lyapunov(T[0] * 0.5, R[0] @ cov_innov @ R[0].T)
    emp = np.cov(st[:, -1].cpu().numpy().T)
    assert np.abs(emp - P_inf).max() < 0.15 * np.abs(P_inf).max()


def test_conditional_simulation_matches_restatement():
    from pymc_statespace_b200.simulation import conditional_simulation

    rng = np.random.default_rng(5)
    args = random_system(rng, 3, 2, 2, 12)
    o = kn.kalman_filter("standard", *args)
    mus = np.stack([o[0][..., 0], o[0][..., 0] * 0.9])       # U = 2 "posterior draws"
    covs = np.stack([o[2], o[2] * 1.1])
    S = 4
    z = rng.normal(size=(2 * S, 12, 3))
    jit = rng.uniform(1e-12, 1e-8, size=2 * S)
    out = conditional_simulation(_dev(mus), _dev(covs), n_simulations=S, z=_dev(z), jitter=_dev(jit)).cpu().numpy()
    assert out.shape == (2 * S, 12, 3)
    for s in (0, 3, 7):
        ref = osim.mvn_draws_blockdiag(mus[s // S], covs[s // S], z[s], jit[s])
        assert rel_err(out[s], ref) < 1e-10
