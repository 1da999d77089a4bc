"""theta -> state-space matrices maps of the reference models, as data ("base + scatter(theta)").

Every reference model's ``update(theta)`` is a sequence of ``set_subtensor`` writes of slices of the flat
parameter vector into constant matrices, optionally followed by the stationary-covariance Lyapunov solve:

* ``BayesianARMA``       reference ``pymc_statespace/models/SARIMAX.py:10-107``
* ``BayesianVARMAX``     reference ``pymc_statespace/models/VARMAX.py:11-150``
* ``BayesianLocalLevel`` reference ``pymc_statespace/models/local_level.py:7-49``

Here that is a ``StateSpaceSpec``: constant base matrices + for each matrix the list
``(theta_index -> flat element index)``.  The GPU pipeline (``logp.py``) applies the maps with
``kfb_scatter_forward`` and transposes them for the gradient with ``kfb_scatter_backward``.
These are host-side descriptors only (numpy, no arithmetic on the hot path).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

MATRICES = ("a0", "P0", "T", "Z", "R", "H", "Q", "c", "d")


@dataclass
class StateSpaceSpec:
    k_states: int
    k_endog: int
    k_posdef: int
    n_theta: int
    base: Dict[str, np.ndarray]
    maps: Dict[str, List[Tuple[int, int]]] = field(default_factory=dict)  # name -> [(theta_idx, flat_idx)], in write order
    stationary_initialization: bool = False
    param_names: Tuple[str, ...] = ()
    param_slices: Dict[str, slice] = field(default_factory=dict)

    def matrices(self, theta: np.ndarray) -> Dict[str, np.ndarray]:
        """Host (numpy) evaluation of the map for ONE theta - used by tests/oracle only."""
        out = {}
        for k in MATRICES:
            mat = np.array(self.base[k], dtype=np.float64, copy=True)
            flat = mat.reshape(-1)
            for ti, fi in self.maps.get(k, []):
                flat[fi] = theta[ti]
            out[k] = mat
        return out


def _empty_base(m, p, r):
    return {"a0": np.zeros((m, 1)), "P0": np.zeros((m, m)), "T": np.zeros((m, m)), "Z": np.zeros((p, m)),
            "R": np.zeros((m, r)), "H": np.zeros((p, p)), "Q": np.zeros((r, r)), "c": np.zeros((m, 1)),
            "d": np.zeros((p, 1))}


def arma_spec(order: Tuple[int, int], stationary_initialization: bool = True) -> StateSpaceSpec:
    """BayesianARMA (reference models/SARIMAX.py:10-107).  theta = [x0 (m), (P0 (m*m)), sigma_state (1),
    rho (p), theta (q)]; ``sigma_state`` is written straight into Q[0,0] (a variance);
    obs_cov stays 0."""
    p_ar, q_ma = order
    m = max(p_ar, q_ma + 1)
    base = _empty_base(m, 1, 1)
    base["Z"][0, 0] = 1.0
    base["T"] = np.eye(m, k=1)
    base["R"][0, 0] = 1.0
    base["P0"] = np.eye(m)
    maps: Dict[str, List[Tuple[int, int]]] = {k: [] for k in MATRICES}
    slices = {}
    cur = 0
    slices["x0"] = slice(cur, cur + m)
    maps["a0"] = [(cur + i, i) for i in range(m)]
    cur += m
    if not stationary_initialization:
        slices["P0"] = slice(cur, cur + m * m)
        maps["P0"] = [(cur + i, i) for i in range(m * m)]
        cur += m * m
    slices["sigma_state"] = slice(cur, cur + 1)
    maps["Q"] = [(cur, 0)]
    cur += 1
    slices["rho"] = slice(cur, cur + p_ar)
    maps["T"] = [(cur + i, i * m + 0) for i in range(p_ar)]  # transition[i, 0]
    cur += p_ar
    slices["theta"] = slice(cur, cur + q_ma)
    maps["R"] = [(cur + i, (i + 1) * 1 + 0) for i in range(q_ma)]  # selection[i+1, 0]
    cur += q_ma
    names = ["x0", "P0", "sigma_state", "rho", "theta"]
    if stationary_initialization:
        names.remove("P0")
    return StateSpaceSpec(m, 1, 1, cur, base, maps, stationary_initialization, tuple(names), slices)


def varmax_spec(k_endog: int, order: Tuple[int, int], stationary_initialization: bool = True,
                measurement_error: bool = True) -> StateSpaceSpec:
    """BayesianVARMAX (reference models/VARMAX.py:11-150).  theta = [x0, (P0), ar_params, ma_params,
    state_cov (k*k, written as a full matrix), (obs_cov diag)]."""
    p_ar, q_ma = order
    k = k_endog
    k_order = max(p_ar, 1) + q_ma
    m = k * k_order
    r = k
    base = _empty_base(m, k, r)
    base["Z"][np.arange(k), np.arange(k)] = 1.0
    if p_ar > 1:
        base["T"][k:k * p_ar, 0:k * (p_ar - 1)] = np.eye(k * (p_ar - 1))
    if q_ma > 1:
        base["T"][m - k * (q_ma - 1):, m - k * q_ma:m - k] = np.eye(k * (q_ma - 1))
    base["R"][0:k, :] = np.eye(k)
    if q_ma > 0:
        start = m - k * q_ma
        base["R"][start:start + k, :] = np.eye(k)
    maps: Dict[str, List[Tuple[int, int]]] = {kk: [] for kk in MATRICES}
    slices = {}
    cur = 0
    slices["x0"] = slice(cur, cur + m)
    maps["a0"] = [(cur + i, i) for i in range(m)]
    cur += m
    if not stationary_initialization:
        slices["P0"] = slice(cur, cur + m * m)
        maps["P0"] = [(cur + i, i) for i in range(m * m)]
        cur += m * m
    if p_ar > 0:
        cnt = k * k * p_ar
        slices["ar_params"] = slice(cur, cur + cnt)
        cols = k * p_ar
        maps["T"] += [(cur + i * cols + j, i * m + j) for i in range(k) for j in range(cols)]
        cur += cnt
    if q_ma > 0:
        cnt = k * k * q_ma
        slices["ma_params"] = slice(cur, cur + cnt)
        c0 = k * max(1, p_ar)
        cols = m - c0
        maps["T"] += [(cur + i * cols + j, i * m + c0 + j) for i in range(k) for j in range(cols)]
        cur += cnt
    slices["state_cov"] = slice(cur, cur + r * r)
    maps["Q"] = [(cur + i, i) for i in range(r * r)]
    cur += r * r
    if measurement_error:
        slices["obs_cov"] = slice(cur, cur + k)
        maps["H"] = [(cur + i, i * k + i) for i in range(k)]
        cur += k
    names = ["x0", "P0", "ar_params", "ma_params", "state_cov", "obs_cov"]
    if stationary_initialization:
        names.remove("P0")
    if not measurement_error:
        names.remove("obs_cov")
    if p_ar == 0:
        names.remove("ar_params")
    if q_ma == 0:
        names.remove("ma_params")
    return StateSpaceSpec(m, k, r, cur, base, maps, stationary_initialization, tuple(names), slices)


def local_level_spec() -> StateSpaceSpec:
    """BayesianLocalLevel (reference models/local_level.py:7-49) - a local LINEAR TREND, k_states = 2
    (SURVEY A.2-Q11).  theta = [x0 (2), P0 (4), sigma_obs (1), sigma_state (2)]."""
    m = r = 2
    base = _empty_base(m, 1, r)
    base["Z"] = np.array([[1.0, 0.0]])
    base["T"] = np.array([[1.0, 1.0], [0.0, 1.0]])
    base["R"] = np.eye(2)
    base["P0"] = np.eye(2)
    maps = {k: [] for k in MATRICES}
    maps["a0"] = [(0, 0), (1, 1)]
    maps["P0"] = [(2 + i, i) for i in range(4)]
    maps["H"] = [(6, 0)]
    maps["Q"] = [(7, 0), (8, 3)]
    slices = {"x0": slice(0, 2), "P0": slice(2, 6), "sigma_obs": slice(6, 7), "sigma_state": slice(7, 9)}
    return StateSpaceSpec(m, 1, r, 9, base, maps, False, ("x0", "P0", "sigma_obs", "sigma_state"), slices)


def local_level_1state_spec() -> StateSpaceSpec:
    """BASELINE.json configs[0]: the TRUE local level model (k_states = 1; SURVEY.md section 8(d) C1(i), A.2-Q11 - the
    reference class BayesianLocalLevel is a 2-state local linear trend, `local_level_spec`).  y_t = mu_t + eps_t,
    mu_t+1 = mu_t + eta_t:  T = Z = R = [[1]];  theta = [a0, P0, sigma2_obs (H), sigma2_level (Q)]."""
    base = _empty_base(1, 1, 1)
    base["T"][0, 0] = base["Z"][0, 0] = base["R"][0, 0] = 1.0
    maps = {k: [] for k in MATRICES}
    maps["a0"], maps["P0"], maps["H"], maps["Q"] = [(0, 0)], [(1, 0)], [(2, 0)], [(3, 0)]
    slices = {"x0": slice(0, 1), "P0": slice(1, 2), "sigma_obs": slice(2, 3), "sigma_state": slice(3, 4)}
    return StateSpaceSpec(1, 1, 1, 4, base, maps, False, ("x0", "P0", "sigma_obs", "sigma_state"), slices)


def custom_spec(k_states: int, k_endog: int, k_posdef: int, n_theta: int, base: Dict[str, np.ndarray],
                maps: Dict[str, List[Tuple[int, int]]], stationary_initialization: bool = False,
                param_names: Tuple[str, ...] = ()) -> StateSpaceSpec:
    """User-defined model (the README.md:67-124 subclass pattern): constant matrices + index maps."""
    full = _empty_base(k_states, k_endog, k_posdef)
    for k, v in base.items():
        v = np.asarray(v, dtype=np.float64)
        full[k] = v.reshape(full[k].shape)
    return StateSpaceSpec(k_states, k_endog, k_posdef, n_theta, full, {k: list(maps.get(k, [])) for k in MATRICES},
                          stationary_initialization, tuple(param_names))


def trend_seasonal_spec(period: int = 29) -> StateSpaceSpec:
    """BASELINE.json config 4: local linear trend + dummy seasonal of `period` (k_states = 2 + period - 1),
    the pattern of reference examples/'Custom SSM - Daily Seasonality.ipynb' scaled up.
    theta = [sigma2_level, sigma2_slope, sigma2_seasonal, sigma2_obs]; a0 = 0, P0 = I fixed."""
    s = period - 1
    m, r = 2 + s, 3
    base = _empty_base(m, 1, r)
    base["T"][0, 0] = base["T"][0, 1] = base["T"][1, 1] = 1.0
    base["T"][2, 2:] = -1.0
    base["T"][3:, 2:-1] = np.eye(s - 1)
    base["Z"][0, 0] = 1.0
    base["Z"][0, 2] = 1.0
    base["R"][0, 0] = base["R"][1, 1] = base["R"][2, 2] = 1.0
    base["P0"] = np.eye(m)
    maps = {k: [] for k in MATRICES}
    maps["Q"] = [(0, 0), (1, 4), (2, 8)]
    maps["H"] = [(3, 0)]
    return StateSpaceSpec(m, 1, r, 4, base, maps, False, ("sigma2_level", "sigma2_slope", "sigma2_seasonal", "sigma2_obs"))


FUSED_K_STATES = tuple(range(1, 9)) + tuple(range(10, 33, 2))  # sizes with fused hot-path kernels in libkfb200.so


def pad_spec(spec: StateSpaceSpec, k_states: int) -> StateSpaceSpec:
    """The same model embedded in ``k_states`` >= spec.k_states states: the extra states have zero rows / columns in
    T, Z, R, c, a0, P0, so they stay identically zero (mean and covariance), never reach an observation and leave logp,
    every per-step output of the original states and every d logp / d theta unchanged - exactly (the padded products only
    add zeros).  Used by ``KalmanLogp`` to run e.g. a 13-state seasonal model (period 12) on the 14-state tensor-core
    kernels instead of the generic run-time-dims kernels."""
    m, m2, p, r = spec.k_states, int(k_states), spec.k_endog, spec.k_posdef
    if m2 < m:
        raise ValueError("cannot pad to fewer states")
    if m2 == m:
        return spec
    shapes = {"a0": (m2, 1), "P0": (m2, m2), "T": (m2, m2), "Z": (p, m2), "R": (m2, r), "H": (p, p), "Q": (r, r),
              "c": (m2, 1), "d": (p, 1)}
    base = {}
    for k in MATRICES:
        old = np.asarray(spec.base[k], dtype=np.float64)
        old = old.reshape(old.shape[0], -1) if old.ndim else old.reshape(1, 1)
        new = np.zeros(shapes[k])
        new[:old.shape[0], :old.shape[1]] = old
        base[k] = new
    ncols = {"a0": (1, 1), "P0": (m, m2), "T": (m, m2), "Z": (m, m2), "R": (r, r), "H": (p, p), "Q": (r, r), "c": (1, 1), "d": (1, 1)}
    maps = {}
    for k in MATRICES:
        old_c, new_c = ncols[k]
        maps[k] = [(ti, (fi // old_c) * new_c + (fi % old_c)) for ti, fi in spec.maps.get(k, [])]
    return StateSpaceSpec(m2, p, r, spec.n_theta, base, maps, spec.stationary_initialization, spec.param_names,
                          dict(spec.param_slices))
