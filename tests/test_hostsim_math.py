"""The device step programs (csrc/kf_core.cuh) compiled for the HOST and checked against the oracle.
This is how the Kalman / adjoint math is verified in the GPU-less build container; the GPU tests repeat the
same comparisons through the real kernels.  (tests/hostsim is test infrastructure, not a product path.)"""
import numpy as np
import pytest

from oracle import kalman_numpy as kn
from oracle import kalman_torch as kt
from tests import hostsim
from tests.helpers import nile_inputs, random_system, rel_err


def _check(kind, args, c=None, d=None, strict=True, static=False, w=None, tol=1e-9):
    ref = kn.kalman_filter(kind, *args, c=c, d=d, strict_reference=strict)
    outs, g, info = hostsim.run(kind, *args, c=c, d=d, strict=strict, static_dims=static, g_ll_obs=w,
                                g_loglik=(0.0 if w is not None else None))
    assert info == 0
    for a, b in zip(outs, ref):
        assert rel_err(a, b) < tol
    _, gt = kt.loglik_and_grads(kind, *args, c=c, d=d, strict_reference=strict, g_ll_obs=w)
    for k in gt:
        assert rel_err(g[k], gt[k]) < tol or np.abs(g[k] - gt[k]).max() < 1e-13, (kind, k)


@pytest.mark.parametrize("static", [True, False], ids=["ThreadCtx", "CoopCtx"])
@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (3, 2, 2), (4, 3, 2)])
@pytest.mark.parametrize("kind", ["standard", "univariate"])
def test_forward_and_adjoint(kind, dims, static):
    m, p, r = dims
    rng = np.random.default_rng(m * 10 + p)
    args = random_system(rng, m, p, r, 25, n_missing=2)
    c, d = rng.normal(size=(m, 1)), rng.normal(size=(p, 1))
    _check(kind, args, c, d, True, static)
    _check(kind, args, c, d, False, static, w=rng.normal(size=25))


@pytest.mark.parametrize("static", [True, False], ids=["ThreadCtx", "CoopCtx"])
def test_single_cholesky_p1_and_steady_state(static):
    rng = np.random.default_rng(2)
    args = random_system(rng, 3, 1, 2, 30, n_missing=3)
    d = rng.normal(size=(1, 1))
    _check("single", args, None, d, True, static)
    _check("cholesky", args, None, d, True, static)
    args = random_system(rng, 4, 2, 2, 30)
    _check("steady_state", args, None, rng.normal(size=(2, 1)), True, static, tol=1e-8)
    _check("steady_state", args, None, rng.normal(size=(2, 1)), False, static, tol=1e-8)


def test_time_varying_and_partial_missing():
    rng = np.random.default_rng(1)
    n, m, p, r = 12, 3, 2, 2
    systems = [random_system(rng, m, p, r, n) for _ in range(n)]
    y, a0, P0 = systems[0][:3]
    T, Z, R, H, Q = (np.stack([s[i] for s in systems]) for i in range(3, 8))
    _check("standard", (y, a0, P0, T, Z, R, H, Q), rng.normal(size=(n, m, 1)), rng.normal(size=(n, p, 1)))
    args = random_system(rng, 4, 3, 2, 30, n_missing=2, partial=True, diag_H=True)
    _check("univariate", args)
    _, _, info = hostsim.run("standard", *args, do_bwd=False)
    assert info < 0


def test_deferred_log_matches_per_step_log():
    args = nile_inputs(5)
    ref = kn.kalman_filter("standard", *args)[4]
    outs, _, _ = hostsim.run("standard", *args, static_dims=True, full=False, do_bwd=False)
    assert abs(outs[4] - ref) < 1e-12 * abs(ref)
