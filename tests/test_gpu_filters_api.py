"""The reference's own filter tests (tests/test_kalman_filter.py), re-run against the B200 filter classes.
The reference compares against statsmodels live; here the comparison target is the oracle, which
tests/test_oracle_pins.py pins on everything available offline."""
import numpy as np
import pytest
import torch

from oracle import kalman_numpy as kn
from oracle import kalman_torch as kt
from tests.helpers import make_test_inputs, nile_inputs, random_system, rel_err

pytestmark = pytest.mark.gpu

output_names = ["filtered_states", "predicted_states", "filtered_covs", "predicted_covs", "log_likelihood", "ll_obs"]


def _filters():
    from pymc_statespace_b200.filters import (CholeskyFilter, SingleTimeseriesFilter, StandardFilter,
                                              SteadyStateFilter, UnivariateFilter)

    return {"StandardFilter": StandardFilter, "CholeskyFilter": CholeskyFilter, "UnivariateFilter": UnivariateFilter,
            "SingleTimeSeriesFilter": SingleTimeseriesFilter, "SteadyStateFilter": SteadyStateFilter}


filter_names = ["StandardFilter", "CholeskyFilter", "UnivariateFilter", "SingleTimeSeriesFilter", "SteadyStateFilter"]


def get_expected_shape(name, p, m, r, n):
    # reference tests/utilities/test_helpers.py:79-90
    if name == "log_likelihood":
        return ()
    if name == "ll_obs":
        return (n,)
    filter_type, variable = name.split("_")
    if filter_type == "predicted":
        n += 1
    return (n, m, 1) if variable == "states" else (n, m, m)


def test_base_class_update_raises():
    from pymc_statespace_b200.filters import BaseFilter

    with pytest.raises(NotImplementedError):
        BaseFilter().update(*[None] * 8)


@pytest.mark.parametrize("filter_name", filter_names)
@pytest.mark.parametrize("dims", [(1, 1, 1, 10), (1, 2, 2, 10), (1, 5, 2, 10), (1, 5, 1, 10)])
def test_output_shapes(filter_name, dims):
    # reference :66-99,159-168
    p, m, r, n = dims
    outputs = _filters()[filter_name]().build_graph(*make_test_inputs(p, m, r, n))
    for name, out in zip(output_names, outputs):
        assert np.shape(out) == get_expected_shape(name, p, m, r, n), name


def test_output_shapes_with_time_varying_matrices():
    # reference :102-156
    from pymc_statespace_b200.filters import StandardFilter

    p, m, r, n = 1, 5, 2, 10
    data, a0, P0, T, Z, R, H, Q = make_test_inputs(p, m, r, n)
    T, Z, R, H, Q = (np.concatenate([np.expand_dims(x, 0)] * n, axis=0) for x in (T, Z, R, H, Q))
    outputs = StandardFilter().build_graph(data, a0, P0, T, Z, R, H, Q)
    static = StandardFilter().build_graph(*make_test_inputs(p, m, r, n))
    for name, out, ref in zip(output_names, outputs, static):
        assert np.shape(out) == get_expected_shape(name, p, m, r, n)
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    with pytest.raises(AssertionError, match="first dimension of a time varying matrix"):
        StandardFilter().build_graph(data, a0, P0, T[:-1], Z, R, H, Q)


@pytest.mark.parametrize("filter_name", filter_names)
def test_output_with_multiple_observed(filter_name):
    # reference :171-189
    p, m, r, n = 5, 5, 1, 10
    inputs = make_test_inputs(p, m, r, n)
    flt = _filters()[filter_name]()
    if filter_name == "SingleTimeSeriesFilter":
        with pytest.raises(AssertionError, match="UnivariateTimeSeries filter requires data be at most 1-dimensional"):
            flt.build_graph(*inputs)
    else:
        outputs = flt.build_graph(*inputs)
        for name, out in zip(output_names, outputs):
            assert np.shape(out) == get_expected_shape(name, p, m, r, n)


@pytest.mark.parametrize("filter_name", filter_names)
@pytest.mark.parametrize("p", [1, 5])
def test_missing_data(filter_name, p):
    # reference :192-209
    m, r, n = 5, 1, 10
    inputs = make_test_inputs(p, m, r, n, missing_data=1)
    flt = _filters()[filter_name]()
    if p > 1 and filter_name == "SingleTimeSeriesFilter":
        with pytest.raises(AssertionError):
            flt.build_graph(*inputs)
    else:
        for out in flt.build_graph(*inputs):
            assert not np.any(np.isnan(out))


@pytest.mark.parametrize("filter_name", filter_names)
@pytest.mark.parametrize("n_missing", [0, 5])
def test_filters_match_oracle_on_nile_fixture(filter_name, n_missing):
    # reference :226-241 (there: vs statsmodels, atol=1e-7, steady-state excluded; here all five vs the oracle)
    if filter_name == "SteadyStateFilter" and n_missing:
        pytest.skip("fixed-gain recursion with missing rows diverges numerically (P0 = 1e6 I) in the reference too")
    kind = {"StandardFilter": "standard", "CholeskyFilter": "cholesky", "UnivariateFilter": "univariate",
            "SingleTimeSeriesFilter": "single", "SteadyStateFilter": "steady_state"}[filter_name]
    inputs = nile_inputs(n_missing)
    outputs = _filters()[filter_name]().build_graph(*inputs)
    ref = kn.kalman_filter(kind, *inputs)
    for name, out, want in zip(output_names, outputs, ref):
        np.testing.assert_allclose(out, want, rtol=1e-8, atol=1e-8 * np.abs(want).max(), err_msg=name)


def test_torch_autograd_through_filter():
    from pymc_statespace_b200.filters import StandardFilter, UnivariateFilter

    rng = np.random.default_rng(8)
    args = random_system(rng, 3, 2, 2, 20, n_missing=2)
    for cls, kind in ((StandardFilter, "standard"), (UnivariateFilter, "univariate")):
        ts = [torch.tensor(a, device="cuda", requires_grad=(i > 0)) for i, a in enumerate(args)]
        outs = cls().build_graph(*ts)
        w = torch.tensor(rng.normal(size=20), device="cuda")
        (outs[4] * 0.7 + (outs[5] * w).sum()).backward()
        _, g1 = kt.loglik_and_grads(kind, *args)
        _, g2 = kt.loglik_and_grads(kind, *args, g_ll_obs=w.cpu().numpy())
        for t, name in zip(ts[1:], ("a0", "P0", "T", "Z", "R", "H", "Q")):
            want = 0.7 * g1[name] + g2[name]
            assert rel_err(t.grad.cpu().numpy(), want) < 1e-8, name


def test_numerical_failures_raise_like_scipy():
    from pymc_statespace_b200.filters import StandardFilter, UnivariateFilter

    rng = np.random.default_rng(9)
    args = list(random_system(rng, 3, 2, 2, 12))
    args[0][4, 0] = np.nan  # partially missing row: reference raises LinAlgError
    with pytest.raises(np.linalg.LinAlgError):
        StandardFilter().build_graph(*args)
    UnivariateFilter().build_graph(*args)  # fine
    args = list(random_system(rng, 3, 2, 2, 12))
    args[6] = -np.eye(2) * 50.0  # H negative definite -> F not PD
    with pytest.raises(np.linalg.LinAlgError):
        StandardFilter().build_graph(*args)


def test_steady_state_rejects_time_varying():
    from pymc_statespace_b200.filters import SteadyStateFilter

    data, a0, P0, T, Z, R, H, Q = make_test_inputs(1, 2, 2, 10)
    with pytest.raises(ValueError, match="time-invariant"):
        SteadyStateFilter().build_graph(data, a0, P0, np.stack([T] * 10), Z, R, H, Q)


@pytest.mark.parametrize("filter_name", filter_names)
def test_last_smoother_is_last_filtered_and_matches_oracle(filter_name):
    # reference tests/test_kalman_filter.py:212-222 + smoothed outputs of :226-241
    from pymc_statespace_b200.filters import KalmanSmoother

    p, m, r, n = 1, 5, 1, 10
    inputs = make_test_inputs(p, m, r, n)
    outs = _filters()[filter_name]().build_graph(*inputs)
    ss, sc = KalmanSmoother().build_graph(inputs[3], inputs[5], inputs[7], outs[0], outs[2])
    assert ss.shape == (n, m, 1) and sc.shape == (n, m, m)
    np.testing.assert_allclose(outs[0][-1], ss[-1])
    np.testing.assert_allclose(outs[2][-1], sc[-1])
    rs, rc = kn.kalman_smoother(inputs[3], inputs[5], inputs[7], outs[0], outs[2])
    np.testing.assert_allclose(ss, rs, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(sc, rc, rtol=1e-8, atol=1e-10)


def test_smoother_nile_and_batched():
    from pymc_statespace_b200 import BatchedKalman, rts_smoother
    from pymc_statespace_b200.filters import KalmanSmoother, StandardFilter

    inputs = nile_inputs(5)
    outs = StandardFilter().build_graph(*inputs)
    ss, sc = KalmanSmoother().build_graph(inputs[3], inputs[5], inputs[7], outs[0], outs[2])
    rs, rc = kn.kalman_smoother(inputs[3], inputs[5], inputs[7], outs[0], outs[2])
    np.testing.assert_allclose(ss, rs, rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(sc[5:], rc[5:], rtol=1e-7, atol=1e-7)  # reference skips the first 5 as well (:236-238)
    # batched, k_states = 6 (warp mode) and 30 (CTA mode)
    rng = np.random.default_rng(2)
    for m, pdim, r, B in ((6, 3, 3, 9), (30, 1, 3, 3)):
        n = 20
        systems = [random_system(rng, m, pdim, r, n, scale_T=0.2) for _ in range(B)]
        y = systems[0][0]
        dev = lambda i: torch.as_tensor(np.stack([s[i] for s in systems]), device="cuda")  # noqa: E731
        bk = BatchedKalman("standard", n, m, pdim, r, n_draws=B)
        out = bk.forward(torch.as_tensor(y[..., 0], device="cuda"), dev(1), dev(2), dev(3), dev(4), dev(5), dev(6), dev(7),
                         outputs=("filtered_states", "filtered_covs"))
        ss, sc = rts_smoother(dev(3), dev(5), dev(7), out["filtered_states"], out["filtered_covs"])
        for b in (0, B - 1):
            o = kn.kalman_filter("standard", y, *systems[b][1:])
            rs, rc = kn.kalman_smoother(systems[b][3], systems[b][5], systems[b][7], o[0], o[2])
            # pinv(P_hat) amplifies rounding by cond(P_hat) (~1e7 for the k_states = 30 system, whose state noise has
            # rank 3): Jacobi-eigen pinv and numpy's SVD pinv agree to ~cond * eps there
            tol = 1e-8 if m < 10 else 1e-6
            assert rel_err(ss[b].cpu().numpy(), rs[..., 0]) < tol
            assert rel_err(sc[b].cpu().numpy(), rc) < tol


def test_pytensor_op_perform_paths_without_pytensor():
    """PyTensor is not installable here, but the numeric halves of the adapter (KalmanFilterOp.perform and
    KalmanFilterGradOp.perform: numpy in, numpy out, argument wiring incl. optional c / d) need no PyTensor at all."""
    from pymc_statespace_b200.pytensor_op import KalmanFilterGradOp, KalmanFilterOp

    rng = np.random.default_rng(17)
    args = random_system(rng, 3, 2, 2, 15, n_missing=2)
    c, d = rng.normal(size=(3, 1)), rng.normal(size=(2, 1))
    for kind, has_c, has_d in (("standard", True, True), ("univariate", False, True), ("cholesky", False, False)):
        strict = kind != "cholesky"
        inputs = list(args) + ([c] if has_c else []) + ([d] if has_d else [])
        op = KalmanFilterOp(kind, strict, has_c, has_d)
        storage = [[None] for _ in range(6)]
        op.perform(None, inputs, storage)
        ref = kn.kalman_filter(kind, *args, c=c if has_c else None, d=d if has_d else None, strict_reference=strict)
        for s, want in zip(storage, ref):
            np.testing.assert_allclose(s[0], want, rtol=1e-8, atol=1e-10)
        assert storage[4][0].shape == () and storage[5][0].shape == (15,)
        w = rng.normal(size=15)
        gop = KalmanFilterGradOp(kind, strict, has_c, has_d)
        gstorage = [[None] for _ in range(len(inputs) - 1)]
        gop.perform(None, inputs + [np.asarray(0.5), w], gstorage)
        _, g1 = kt.loglik_and_grads(kind, *args, c=c if has_c else None, d=d if has_d else None, strict_reference=strict)
        _, g2 = kt.loglik_and_grads(kind, *args, c=c if has_c else None, d=d if has_d else None, strict_reference=strict,
                                    g_ll_obs=w)
        names = ["a0", "P0", "T", "Z", "R", "H", "Q"] + (["c"] if has_c else []) + (["d"] if has_d else [])
        for s, name, x in zip(gstorage, names, inputs[1:]):
            want = 0.5 * g1[name] + g2[name]
            assert s[0].shape == np.shape(x)
            if kind == "cholesky" and name in ("P0", "H"):
                continue  # gauge of the symmetric inputs differs for the Cholesky-based filter
            assert rel_err(s[0], want) < 1e-8, (kind, name)


def test_seam_graph_is_reused_and_matches_oracle():
    """numpy in / numpy out at one model per call (seam.SeamGraph): the cached CUDA graph of a geometry is replayed with
    new values, with and without a per-step cotangent, with c / d and time-varying T; failures still raise."""
    from pymc_statespace_b200 import seam
    from pymc_statespace_b200.filters import StandardFilter, UnivariateFilter

    rng = np.random.default_rng(11)
    n, m, p, r = 30, 3, 2, 2
    flt = StandardFilter()
    seam._CACHE.clear()
    for rep in range(3):
        y, a0, P0, T, Z, R, H, Q = random_system(rng, m, p, r, n, n_missing=2)
        c, d = rng.normal(size=(m, 1)), rng.normal(size=(p, 1))
        arrays = {"data": y, "a0": a0, "P0": P0, "T": T, "Z": Z, "R": R, "H": H, "Q": Q, "c": c, "d": d}
        outs = flt.build_graph(y, a0, P0, T, Z, R, H, Q, c=c, d=d)
        ref = kn.kalman_filter("standard", y, a0, P0, T, Z, R, H, Q, c=c, d=d)
        for o, e in zip(outs, ref):
            assert rel_err(np.asarray(o), np.asarray(e)) < 1e-9
        w, wt = float(rng.normal()), rng.normal(size=n)
        _, g1 = kt.loglik_and_grads("standard", y, a0, P0, T, Z, R, H, Q, c=c, d=d)
        _, g2 = kt.loglik_and_grads("standard", y, a0, P0, T, Z, R, H, Q, c=c, d=d, g_ll_obs=wt)
        for g_obs in (None, wt):
            ll, g = seam.logp_grads_numpy(flt, arrays, g_loglik=w, g_ll_obs=g_obs)
            assert abs(ll - float(ref[4])) < 1e-9 * abs(float(ref[4]))
            for k in ("a0", "P0", "T", "Z", "R", "H", "Q", "c", "d"):
                want = w * g1[k] + (0.0 if g_obs is None else g2[k])
                assert g[k].shape == arrays[k].shape
                assert rel_err(g[k], np.asarray(want).reshape(g[k].shape)) < 1e-8, (k, rep, g_obs is None)
    assert len(seam._CACHE) == 3  # six outputs ; grad ; grad with g_ll_obs - each captured once, replayed three times
    # time-varying T (time-first, filters/utilities.py:9-14) is its own geometry
    y, a0, P0, T, Z, R, H, Q = random_system(rng, m, p, r, n)
    Tt = np.repeat(T[None], n, axis=0) * (1.0 + 0.05 * rng.normal(size=(n, 1, 1)))
    ll, g = seam.logp_grads_numpy(flt, {"data": y, "a0": a0, "P0": P0, "T": Tt, "Z": Z, "R": R, "H": H, "Q": Q})
    lref, gref = kt.loglik_and_grads("standard", y, a0, P0, Tt, Z, R, H, Q)
    assert abs(ll - lref) < 1e-9 * abs(lref) and g["T"].shape == Tt.shape
    assert rel_err(g["T"], np.asarray(gref["T"])) < 1e-8
    # a partially missing row: LinAlgError for the standard filter, fine for the univariate one
    yb = y.copy()
    yb[4, 0, 0] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        flt.build_graph(yb, a0, P0, T, Z, R, H, Q)
    UnivariateFilter().build_graph(yb, a0, P0, T, Z, R, H, Q)


@pytest.mark.parametrize("m", [2, 3, 6])
def test_smoother_singular_predicted_covariance_uses_pinv(m):
    """RTS smoother with a SINGULAR P_hat (a state without noise and without initial uncertainty): pinv(P_hat) differs
    from any inverse, so the kernels must take the eigendecomposition path with numpy's cutoff (kalman_smoother.py:92) and
    not the Cholesky fast path; regular units of the same batch take the fast path.  k_states 2, 3: thread per unit;
    6: one warp per unit."""
    from pymc_statespace_b200 import rts_smoother

    rng = np.random.default_rng(40 + m)
    B, n, r = 4, 12, 1
    fs = rng.normal(size=(B, n, m))
    A = rng.normal(size=(B, n, m, m))
    fc = A @ np.swapaxes(A, -1, -2) + 0.1 * np.eye(m)
    T = np.tile(np.eye(m), (B, 1, 1)) * 0.9
    R = np.zeros((B, m, r))
    R[:, 0, 0] = 1.0                      # only state 0 is driven by noise
    Q = np.tile(np.array([[0.5]]), (B, 1, 1))
    # units 1 and 3: the last state is known exactly at every step -> P_hat has a zero row / column
    for b in (1, 3):
        fc[b, :, m - 1, :] = 0.0
        fc[b, :, :, m - 1] = 0.0
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")  # noqa: E731
    ss, sc = rts_smoother(dev(T), dev(R), dev(Q), dev(fs), dev(fc))
    for b in range(B):
        rs, rc = kn.kalman_smoother(T[b], R[b], Q[b], fs[b][..., None], fc[b])
        assert rel_err(ss[b].cpu().numpy(), rs[..., 0]) < 1e-8, b
        assert rel_err(sc[b].cpu().numpy(), rc) < 1e-8, b
