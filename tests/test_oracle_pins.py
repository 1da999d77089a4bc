"""Pins the CPU oracle on everything that is available offline.

The reference cannot be imported here (PyTensor / PyMC / statsmodels absent) and its own tests hold no stored
golden vectors (they call statsmodels live), so the oracle is pinned on:
  * an algorithm-independent identity: Kalman loglik == dense multivariate-normal log-density of the stacked sample;
  * cross-filter identities implied by reference tests/test_kalman_filter.py:226-241 passing for four filters;
  * the DARE known-answer test of reference tests/test_pytensor_scipy.py:55-73 (scipy "darex #1");
  * the soft known answer llf = 1950.186 of the VAR(2) printed in reference examples/'VARMAX Example.ipynb':363;
  * the documented quirk offsets;
  * committed golden vectors of the oracle itself (tests/golden/*.npz, made by tests/golden/make_golden.py) so that
    any later change of the oracle is caught.
  * GRADIENTS: torch-autograd of the dense density (oracle.kalman_torch.dense_gaussian_loglik) - an algorithm-independent
    known answer for all nine matrix cotangents and for d logp / d theta of BayesianARMA.
Everything remains "parity unpinned" against the real reference itself (it has never run next to this code).
"""
import os

import numpy as np
import pytest
import scipy.linalg

from oracle import kalman_numpy as kn
from oracle import kalman_torch as kt
from tests.helpers import GOLDEN, make_test_inputs, nile_data, nile_inputs, random_system, rel_err


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (3, 2, 2), (4, 3, 2)])
def test_loglik_equals_dense_gaussian_density(dims):
    m, p, r = dims
    args = random_system(np.random.default_rng(m * 7 + p), m, p, r, 14)
    dense = kn.dense_gaussian_loglik(*args)
    assert abs(kn.kalman_filter("standard", *args, strict_reference=False)[4] - dense) < 1e-10 * abs(dense)
    assert abs(kn.kalman_filter("cholesky", *args, strict_reference=False)[4] - dense) < 1e-10 * abs(dense)
    if p == 1:
        for kind in ("standard", "cholesky", "single", "univariate"):
            assert abs(kn.kalman_filter(kind, *args)[4] - dense) < 1e-10 * abs(dense)


def test_univariate_equals_dense_density_for_diagonal_H():
    args = random_system(np.random.default_rng(5), 4, 3, 2, 12, diag_H=True)
    dense = kn.dense_gaussian_loglik(*args)
    assert abs(kn.kalman_filter("univariate", *args)[4] - dense) < 1e-10 * abs(dense)


@pytest.mark.parametrize("n_missing", [0, 5])
def test_four_filters_agree_on_nile_fixture(n_missing):
    # reference tests/test_kalman_filter.py:226-241: all four match statsmodels => they match each other
    args = nile_inputs(n_missing)
    ref = kn.kalman_filter("standard", *args)
    for kind in ("cholesky", "single", "univariate"):
        out = kn.kalman_filter(kind, *args)
        for a, b in zip(out, ref):
            np.testing.assert_allclose(a, b, rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("n_missing", [0, 5])
def test_nile_fixture_equals_dense_gaussian_density(n_missing):
    """the one fixture the reference value-tests (Nile local linear trend, P0 = 1e6 I,
    tests/test_kalman_filter.py:226-241) pinned on an algorithm-independent known answer - the dense multivariate-normal
    log-density of the stacked sample (missing rows deleted), evaluated in 40-digit arithmetic and committed as
    tests/golden/nile_dense_loglik.json (tests/golden/make_nile_dense.py).  All four filters the reference compares."""
    import json

    gold = json.load(open(os.path.join(GOLDEN, "nile_dense_loglik.json")))[str(n_missing)]
    args = nile_inputs(n_missing)
    assert sorted(np.where(np.isnan(args[0][:, 0, 0]))[0].tolist()) == gold["missing_rows"]
    for kind in ("standard", "cholesky", "single", "univariate"):
        ll = kn.kalman_filter(kind, *args)[4]
        assert abs(ll - gold["loglik"]) < 1e-12 * abs(gold["loglik"]), (kind, ll, gold["loglik"])


def test_dense_density_generator_reproduces_golden_value():
    # the committed golden value is what the generator produces today (40-digit dense density, ~3 s)
    import json

    gold = json.load(open(os.path.join(GOLDEN, "nile_dense_loglik.json")))["5"]
    assert kn.dense_gaussian_loglik(*nile_inputs(5), mp_digits=40) == gold["loglik"]


def test_dense_density_with_partial_missing_rows():
    # univariate filter (the only one that supports partially missing rows) against the dense density, float64 and mp
    args = random_system(np.random.default_rng(3), 4, 3, 2, 12, n_missing=2, partial=True, diag_H=True)
    ll = kn.kalman_filter("univariate", *args)[4]
    assert abs(ll - kn.dense_gaussian_loglik(*args)) < 1e-10 * abs(ll)
    assert abs(ll - kn.dense_gaussian_loglik(*args, mp_digits=30)) < 1e-12 * abs(ll)


def test_standard_filter_log2pi_quirk_offset():
    # strict standard filter counts log(2 pi) once per step instead of k_endog times
    p, n = 3, 20
    args = random_system(np.random.default_rng(9), 6, p, 3, n)
    strict = kn.kalman_filter("standard", *args, strict_reference=True)[4]
    fixed = kn.kalman_filter("standard", *args, strict_reference=False)[4]
    assert abs((strict - fixed) - 0.5 * (p - 1) * n * kn.LOG_2PI) < 1e-9


def test_strict_cholesky_is_exact_only_for_p1():
    a1 = random_system(np.random.default_rng(1), 3, 1, 1, 15)
    assert abs(kn.kalman_filter("cholesky", *a1)[4] - kn.kalman_filter("standard", *a1)[4]) < 1e-10
    a3 = random_system(np.random.default_rng(2), 4, 3, 2, 15)
    strict = kn.kalman_filter("cholesky", *a3, strict_reference=True)[4]
    fixed = kn.kalman_filter("cholesky", *a3, strict_reference=False)[4]
    assert abs(strict - fixed) > 1e-3  # A.2-Q4: the as-coded filter is wrong for p > 1


def test_dare_known_answer_darex1():
    # reference tests/test_pytensor_scipy.py:55-73
    a = np.array([[4.0, 3.0], [-4.5, -3.5]])
    b = np.array([[1.0], [-1.0]])
    q = np.array([[9.0, 6.0], [6.0, 4.0]])
    r = np.array([[1.0]])
    import torch

    x = kt.solve_discrete_are(*(torch.tensor(v) for v in (a, b, q, r))).numpy()
    res = a.T @ x @ a - x - (a.T @ x @ b) @ np.linalg.solve(r + b.T @ x @ b, b.T @ x @ a) + q
    np.testing.assert_allclose(res, np.zeros_like(res), atol=1e-12)


def test_dare_and_lyapunov_adjoints_match_finite_differences():
    import torch

    rng = np.random.default_rng(3)
    m, p = 3, 2
    A = rng.normal(size=(m, m)) * 0.4
    B = rng.normal(size=(m, p))
    L = rng.normal(size=(m, m)); Q = L @ L.T + np.eye(m)
    L = rng.normal(size=(p, p)); R = L @ L.T + np.eye(p)
    W = rng.normal(size=(m, m))
    ts = [torch.tensor(v, requires_grad=True) for v in (A, B, Q, R)]
    (kt.solve_discrete_are(*ts) * torch.tensor(W)).sum().backward()

    def f(A_, B_, Q_, R_):
        return float((scipy.linalg.solve_discrete_are(A_, B_, 0.5 * (Q_ + Q_.T), 0.5 * (R_ + R_.T)) * W).sum())

    eps = 1e-6
    for idx, (M, t) in enumerate(zip((A, B), ts[:2])):
        fd = np.zeros_like(M)
        for i in range(M.shape[0]):
            for j in range(M.shape[1]):
                d = np.zeros_like(M); d[i, j] = eps
                args_p = [A, B, Q, R]; args_m = [A, B, Q, R]
                args_p[idx] = M + d; args_m[idx] = M - d
                fd[i, j] = (f(*args_p) - f(*args_m)) / (2 * eps)
        np.testing.assert_allclose(t.grad.numpy(), fd, rtol=1e-5, atol=1e-7)
    ts = [torch.tensor(v, requires_grad=True) for v in (A, Q)]
    (kt.solve_discrete_lyapunov(*ts) * torch.tensor(W)).sum().backward()
    fd = np.zeros_like(A)
    for i in range(m):
        for j in range(m):
            d = np.zeros_like(A); d[i, j] = eps
            fd[i, j] = ((scipy.linalg.solve_discrete_lyapunov(A + d, Q) - scipy.linalg.solve_discrete_lyapunov(A - d, Q)) * W).sum() / (2 * eps)
    np.testing.assert_allclose(ts[0].grad.numpy(), fd, rtol=1e-5, atol=1e-7)


def test_var2_macrodata_soft_known_answer():
    # reference examples/'VARMAX Example.ipynb' cell 7: statsmodels VAR(2) MLE llf = 1950.186 with the 4-decimal
    # coefficients printed there.  Rounded coefficients => agreement to ~0.05 (1950.1451).
    path = os.path.join(GOLDEN, "var2_macrodata_params.npz")
    if not os.path.exists(path):
        pytest.skip("coefficient fixture not committed")
    z = np.load(path)
    data = np.loadtxt(os.path.join(GOLDEN, "statsmodels_macrodata_processed.csv"), delimiter=",", skiprows=1, usecols=(1, 2, 3))
    k = 3
    T = np.zeros((6, 6)); T[:3, :3] = z["A1"]; T[:3, 3:] = z["A2"]; T[3:, :3] = np.eye(3)
    R = np.zeros((6, 3)); R[:3] = np.eye(3)
    Zm = np.zeros((3, 6)); Zm[:, :3] = np.eye(3)
    Q = z["L"] @ z["L"].T
    P0 = scipy.linalg.solve_discrete_lyapunov(T, R @ Q @ R.T)
    y = data - z["intercept"][None, :] if "intercept" in z else data
    out = kn.kalman_filter("standard", y[:, :, None], np.zeros((6, 1)), P0, T, Zm, R, np.zeros((k, k)), Q,
                           strict_reference=False)
    assert abs(out[4] - 1950.186) < 0.1


def test_numpy_and_torch_twins_agree():
    rng = np.random.default_rng(0)
    args = random_system(rng, 4, 2, 2, 25, n_missing=2)
    for kind in ("standard", "cholesky", "univariate"):
        for strict in (True, False):
            o = kn.kalman_filter(kind, *args, strict_reference=strict)
            t = kt.kalman_filter(kind, *args, strict_reference=strict)
            for a, b in zip(o, t):
                assert rel_err(np.asarray(b.detach()), a) < 1e-12
    args = random_system(rng, 4, 2, 2, 25)
    o = kn.kalman_filter("steady_state", *args)
    t = kt.kalman_filter("steady_state", *args)
    for a, b in zip(o, t):
        assert rel_err(np.asarray(b.detach()), a) < 1e-12


def test_torch_gradient_matches_finite_differences():
    rng = np.random.default_rng(4)
    args = list(random_system(rng, 3, 2, 2, 12, n_missing=1))
    _, g = kt.loglik_and_grads("standard", *args)
    eps = 1e-6
    for name, idx in (("T", 3), ("Z", 4), ("a0", 1)):
        M = args[idx]
        fd = np.zeros_like(M)
        for i in range(M.shape[0]):
            for j in range(M.shape[1]):
                d = np.zeros_like(M); d[i, j] = eps
                ap = list(args); am = list(args)
                ap[idx] = M + d; am[idx] = M - d
                fd[i, j] = (kn.kalman_filter("standard", *ap)[4] - kn.kalman_filter("standard", *am)[4]) / (2 * eps)
        np.testing.assert_allclose(g[name], fd, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("p,m,r,n", [(1, 1, 1, 10), (1, 2, 2, 10), (1, 5, 2, 10), (1, 5, 1, 10), (5, 5, 1, 10)])
@pytest.mark.parametrize("kind", kn.FILTER_KINDS)
def test_output_shapes_match_reference_table(kind, p, m, r, n):
    # reference tests/test_kalman_filter.py:66-99,159-189 + tests/utilities/test_helpers.py:79-90
    args = make_test_inputs(p, m, r, n)
    if kind == "single" and p > 1:
        with pytest.raises(AssertionError, match="UnivariateTimeSeries filter requires data be at most 1-dimensional"):
            kn.kalman_filter(kind, *args)
        return
    fs, ps, fc, pc, ll, llo = kn.kalman_filter(kind, *args)
    assert fs.shape == (n, m, 1) and ps.shape == (n + 1, m, 1)
    assert fc.shape == (n, m, m) and pc.shape == (n + 1, m, m)
    assert np.ndim(ll) == 0 and llo.shape == (n,)


def test_golden_vectors_of_the_oracle():
    path = os.path.join(GOLDEN, "oracle_golden.npz")
    z = np.load(path)
    for key in [k[:-3] for k in z.files if k.endswith("_ll")]:
        kind, seed, m, p, r, n, miss = key.split("-")
        args = random_system(np.random.default_rng(int(seed)), int(m), int(p), int(r), int(n), n_missing=int(miss))
        out = kn.kalman_filter(kind, *args)
        np.testing.assert_allclose(out[4], z[key + "_ll"], rtol=1e-12)
        np.testing.assert_allclose(out[0], z[key + "_fs"], rtol=1e-10, atol=1e-12)
        _, g = kt.loglik_and_grads(kind, *args)
        np.testing.assert_allclose(g["T"], z[key + "_gT"], rtol=1e-9, atol=1e-12)


def _sym_if_square_sym_input(name, g):
    # P0, H, Q are mathematically symmetric inputs: the entry-wise split of their gradient depends on how a graph uses
    # each entry (DESIGN.md "gradient gauge"); the symmetric part is graph-independent
    return 0.5 * (g + g.T) if name in ("P0", "H", "Q") else g


@pytest.mark.parametrize("n_missing", [0, 3])
@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (3, 2, 2), (4, 3, 2)])
def test_gradient_oracle_equals_autograd_of_dense_density(dims, n_missing):
    """Pins the GRADIENT oracle on an algorithm-independent known answer: torch-autograd of the dense multivariate-normal
    log-density of the stacked sample (oracle.kalman_torch.dense_gaussian_loglik - no recursion, no filter) against
    torch-autograd of the restated filters, for all nine inputs (a0, P0, T, Z, R, H, Q, c, d), with all-missing rows.
    The reference's gradient is autodiff of the same scalar, so this is the value it must have."""
    m, p, r = dims
    rng = np.random.default_rng(900 + 10 * m + p + n_missing)
    args = random_system(rng, m, p, r, 14, n_missing=n_missing)
    c, d = rng.normal(size=(m, 1)), rng.normal(size=(p, 1))
    ll, gd = kt.dense_loglik_and_grads(*args, c=c, d=d)
    assert abs(ll - kn.dense_gaussian_loglik(*args, c=c, d=d)) < 1e-11 * abs(ll)
    kinds = ("standard", "cholesky", "single", "univariate") if p == 1 else ("standard", "cholesky")
    for kind in kinds:
        l2, g2 = kt.loglik_and_grads(kind, *args, c=c, d=d, strict_reference=False)
        assert abs(l2 - ll) < 1e-11 * abs(ll), kind
        for k in gd:
            a, b = _sym_if_square_sym_input(k, g2[k]), _sym_if_square_sym_input(k, gd[k])
            assert rel_err(a, b) < 1e-9, (kind, k)
    if p > 1:  # the univariate filter needs a diagonal H to be the same model
        args = random_system(rng, m, p, r, 14, n_missing=n_missing, diag_H=True)
        ll, gd = kt.dense_loglik_and_grads(*args, c=c, d=d)
        l2, g2 = kt.loglik_and_grads("univariate", *args, c=c, d=d)
        assert abs(l2 - ll) < 1e-11 * abs(ll)
        for k in gd:
            a, b = _sym_if_square_sym_input(k, g2[k]), _sym_if_square_sym_input(k, gd[k])
            if k == "H":  # the univariate filter reads diag(H) only
                a, b = np.diag(a), np.diag(b)
            assert rel_err(a, b) < 1e-9, ("univariate", k)


@pytest.mark.parametrize("order", [(1, 1), (2, 1), (3, 2)])
def test_theta_gradient_of_arma_equals_dense_density(order):
    """theta-level pin for the north-star models (BayesianARMA, stationary initialisation, models/SARIMAX.py:59-107):
    d logp / d theta from the restated model + filter + Lyapunov adjoint (the formula the reference uses) against
    autograd of [theta -> matrices -> P0 by ONE dense Kronecker solve -> dense log-density]; and the plain-C port
    (the timed CPU arm) against the same number."""
    import torch

    from oracle import kalman_c, models as om

    p_, q_ = order
    k_states = max(p_, q_ + 1)
    rng = np.random.default_rng(40 + 10 * p_ + q_)
    n = 40
    y = rng.normal(size=(n, 1, 1))
    theta = np.r_[rng.normal(size=k_states) * 0.3, 0.7, rng.uniform(-0.4, 0.4, size=p_) / np.arange(1, p_ + 1),
                  rng.uniform(-0.5, 0.5, size=q_)]
    fn = lambda th: om.arma_matrices(th, order)  # noqa: E731
    lp, g = om.logp_and_grad_theta(fn, theta, y)

    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    a0, _, T, Z, R, H, Q = om.arma_matrices(th, order)
    P0 = kt.lyapunov_dense(T, R @ Q @ R.T)
    ll = kt.dense_gaussian_loglik(y, a0, P0, T, Z, R, H, Q)
    (gd,) = torch.autograd.grad(ll, [th])
    assert abs(lp - float(ll.detach())) < 1e-11 * abs(lp)
    assert rel_err(g, gd.numpy()) < 1e-9

    # the C port evaluates matrices -> (loglik, matrix cotangents); chain them to theta with the dense map
    th2 = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    a0, _, T, Z, R, H, Q = om.arma_matrices(th2, order)
    RQR = R @ Q @ R.T
    P0 = kt.lyapunov_dense(T, RQR)
    npy = lambda v: v.detach().numpy()  # noqa: E731
    ll_c, gm, bad = kalman_c.logp_grad_batch(y[:, :, 0], npy(a0)[None, :, 0], npy(P0)[None], npy(T)[None], npy(Z), npy(H),
                                             npy(RQR)[None], nthreads=1)
    assert bad == 0 and abs(ll_c[0] - float(ll.detach())) < 1e-11 * abs(lp)
    surrogate = sum((torch.as_tensor(gm[k][0]).reshape(v.shape) * v).sum()
                    for k, v in (("a0", a0), ("P0", P0), ("T", T), ("Z", Z), ("H", H), ("C", RQR)))
    (gc,) = torch.autograd.grad(surrogate, [th2])
    assert rel_err(gc.numpy(), gd.numpy()) < 1e-9


@pytest.mark.parametrize("kind", ["standard", "univariate", "cholesky"])
def test_theta_gradient_of_varmax_equals_dense_density(kind):
    """BASELINE configs[2] at theta level (BayesianVARMAX(2,0), k_endog 3, measurement error, stationary initialisation,
    whole rows missing; models/VARMAX.py:95-150): the oracle the CUDA kernels are compared with in
    tests/test_gpu_logp_theta.py::test_varmax_theta_gradient_with_missing_rows, against autograd of
    [theta -> matrices -> P0 by one dense Kronecker solve -> dense log-density with the missing rows deleted].
    state_cov enters as theta.reshape(3, 3): entry-wise gauge, so its block is compared after symmetrisation."""
    import torch

    from oracle import models as om
    from pymc_statespace_b200.synthetic import varmax20_workload

    spec, y, theta = varmax20_workload(n_draws=3, n=30)
    assert np.isnan(y).any()
    sl = spec.param_slices["state_cov"]
    for b in range(3):
        lp, g = om.logp_and_grad_theta(lambda t: om.varmax_matrices(t, 3, (2, 0), True, True), theta[b], y[:, :, None], kind,
                                       kind != "cholesky")
        th = torch.tensor(theta[b], dtype=torch.float64, requires_grad=True)
        a0, _, T, Z, R, H, Q = om.varmax_matrices(th, 3, (2, 0), True, True)
        ll = kt.dense_gaussian_loglik(y[:, :, None], a0, kt.lyapunov_dense(T, R @ Q @ R.T), T, Z, R, H, Q)
        (gd,) = torch.autograd.grad(ll, [th])
        gd = gd.numpy().copy()
        ll = float(ll.detach())
        if kind == "standard":  # as-coded constant: log(2 pi) x 1 instead of x k_endog per observed row
            ll += 0.5 * (3 - 1) * np.log(2 * np.pi) * int((~np.isnan(y).any(axis=1)).sum())
        assert abs(lp - ll) < 1e-10 * abs(ll), (kind, lp, ll)
        for arr in (g, gd):
            blk = arr[sl].reshape(3, 3)
            arr[sl] = (0.5 * (blk + blk.T)).reshape(-1)
        assert rel_err(g, gd) < 1e-8, kind


def test_nile_local_level_soft_known_answer_from_the_literature():
    """SOFT pin (3 decimals), BASELINE configs[0] (true local level, k_states 1): the Nile local-level model at its
    maximum-likelihood variances is a textbook example (Durbin & Koopman 2012, ch. 2: sigma2_eps = 15,099,
    sigma2_eta = 1,469.1); statsmodels' UnobservedComponents("local level") documentation prints llf = -632.538 for it
    (sigma2.irregular 1.508e+04, sigma2.level 1478.81; approximate diffuse start P0 = 1e6, first observation burned).
    The value is QUOTED FROM MEMORY of public documentation - there is no network here and it is not in /root/reference -
    hence soft: 1e-3 absolute, both parameter sets (the likelihood is flat at its maximum)."""
    y = np.asarray(nile_data(), dtype=float).reshape(-1, 1, 1)
    one = np.eye(1)
    for H, Q in ((15080.0, 1478.81), (15099.0, 1469.1)):
        for kind in ("standard", "cholesky", "single", "univariate"):
            out = kn.kalman_filter(kind, y, np.zeros((1, 1)), 1e6 * one, one, one, one, H * one, Q * one)
            assert abs(out[5][1:].sum() - (-632.538)) < 1e-3, (kind, H, Q, out[5][1:].sum())


@pytest.mark.parametrize("dims", [(2, 1, 1), (3, 2, 2), (4, 3, 2)])
def test_filtered_predicted_and_smoothed_moments_equal_dense_conditional_moments(dims):
    """Pins the per-step outputs (rows a1 / f2) on an algorithm-independent known answer: the conditional moments of the
    joint Gaussian of states and observations by dense linear algebra (oracle.kalman_numpy.dense_gaussian_state_moments):
    filtered = E / Cov[x_t | y_0..t], predicted = [x_{t+1} | y_0..t], RTS-smoothed = [x_t | y_0..n-1]; whole rows missing,
    intercepts c and d (the reference's smoother has no intercept in its predict step, kalman_smoother.py:99-104: c = 0
    there).  All exact filters; univariate with a diagonal H."""
    m, p, r = dims
    n = 10
    rng = np.random.default_rng(60 + m)
    for diag_H in (False, True):
        args = random_system(rng, m, p, r, n, n_missing=2, diag_H=diag_H)
        c, d = rng.normal(size=(m, 1)), rng.normal(size=(p, 1))
        kinds = ("standard", "cholesky") + (("univariate",) if diag_H or p == 1 else ()) + (("single",) if p == 1 else ())
        cond = kn.dense_gaussian_state_moments(*args, c=c, d=d)
        for kind in kinds:
            fs, ps, fc, pc = kn.kalman_filter(kind, *args, c=c, d=d, strict_reference=False)[:4]
            np.testing.assert_allclose(ps[0], args[1])
            np.testing.assert_allclose(pc[0], args[2])
            for t in range(n):
                for (got_a, got_P), (a, P) in (((fs[t], fc[t]), cond(t, t)), ((ps[t + 1], pc[t + 1]), cond(t + 1, t))):
                    assert rel_err(got_a, a) < 1e-10 and rel_err(got_P, P) < 1e-10, (kind, t)
        cond = kn.dense_gaussian_state_moments(*args, d=d)
        out = kn.kalman_filter("standard", *args, d=d, strict_reference=False)
        ss, sc = kn.kalman_smoother(args[3], args[5], args[7], out[0], out[2])
        for t in range(n):
            a, P = cond(t, n - 1)
            assert rel_err(ss[t], a) < 1e-10 and rel_err(sc[t], P) < 1e-10, t


def test_time_varying_loglik_equals_dense_gaussian_density():
    """Row f3: the reference tests time-varying inputs for shapes only (tests/test_kalman_filter.py:102-156).  Values are
    pinned here on the dense density with time-varying T, Z, R, H, Q, c, d (time first, filters/utilities.py:9-14:
    matrices of step t in the update AND the predict of step t), all varying / only T varying / single series."""
    rng = np.random.default_rng(1)
    n, m, p, r = 12, 3, 2, 2
    systems = [random_system(rng, m, p, r, n, n_missing=2) for _ in range(n)]
    y, a0, P0 = systems[0][:3]
    T, Z, R, H, Q = (np.stack([s[i] for s in systems]) for i in range(3, 8))
    c, d = rng.normal(size=(n, m, 1)), rng.normal(size=(n, p, 1))
    dense = kn.dense_gaussian_loglik_time_varying(y, a0, P0, T, Z, R, H, Q, c, d)
    for kind in ("standard", "cholesky"):
        assert abs(kn.kalman_filter(kind, y, a0, P0, T, Z, R, H, Q, c=c, d=d, strict_reference=False)[4] - dense) < 1e-10 * abs(dense)
    dense = kn.dense_gaussian_loglik_time_varying(y, a0, P0, T, Z[0], R[0], H[0], Q[0])
    assert abs(kn.kalman_filter("standard", y, a0, P0, T, Z[0], R[0], H[0], Q[0], strict_reference=False)[4] - dense) < 1e-10 * abs(dense)
    assert abs(kn.dense_gaussian_loglik_time_varying(*systems[0]) - kn.dense_gaussian_loglik(*systems[0])) < 1e-12
    y1 = y[:, :1]
    dense = kn.dense_gaussian_loglik_time_varying(y1, a0, P0, T, Z[:, :1], R, H[:, :1, :1], Q, c, d[:, :1])
    for kind in ("standard", "cholesky", "single"):
        got = kn.kalman_filter(kind, y1, a0, P0, T, Z[:, :1], R, H[:, :1, :1], Q, c=c, d=d[:, :1], strict_reference=False)[4]
        assert abs(got - dense) < 1e-10 * abs(dense), kind


@pytest.mark.parametrize("dims", [(2, 1, 1), (4, 2, 2), (4, 3, 2)])
def test_steady_state_filter_equals_dense_density_started_at_the_riccati_fixed_point(dims):
    """Row a8: on complete data the steady-state filter (kalman_filter.py:354-441) IS the exact filter started at
    P0 = P_ss (P_t = P_ss for every t), so its loglik equals the dense density with that P0 and its gradient equals
    autograd of [matrices -> P_ss by unrolled Riccati iteration -> dense density]: pins the filter, scipy's DARE and
    the DARE adjoint formula (utils/pytensor_scipy.py:39-60) on an answer that shares nothing with them; P0 gets zero
    gradient.  Corrected constants (strict_reference=False: log 2 pi x k_endog)."""
    import torch

    m, p, r = dims
    args = list(random_system(np.random.default_rng(9 + m + p), m, p, r, 14))
    ll, g = kt.loglik_and_grads("steady_state", *args, strict_reference=False)
    names = ("a0", "P0", "T", "Z", "R", "H", "Q")
    ins = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in zip(names, args[1:])}
    T, Z, R, H, Q = (ins[k] for k in ("T", "Z", "R", "H", "Q"))
    Pss = kt.dare_by_riccati_iteration(T, Z, R @ Q @ R.T, H)
    lld = kt.dense_gaussian_loglik(args[0], ins["a0"], Pss, T, Z, R, H, Q)
    gs = torch.autograd.grad(lld, [ins[k] for k in names], allow_unused=True)
    assert abs(ll - float(lld.detach())) < 1e-11 * abs(ll)
    assert np.abs(g["P0"]).max() == 0.0 and gs[1] is None
    for k, gd in zip(names, gs):
        if gd is not None:
            assert rel_err(_sym_if_square_sym_input(k, g[k]), _sym_if_square_sym_input(k, gd.numpy())) < 1e-9, k
