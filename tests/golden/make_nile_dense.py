"""Generates tests/golden/nile_dense_loglik.json: the log-density of the Nile local-linear-trend fixture
(reference tests/utilities/test_helpers.py:110-118: a0 = 0, P0 = 1e6 I, Q = diag(0.5, 0.01), H = 0.8; 0 and 5 whole
rows missing, seed 0) computed WITHOUT any Kalman recursion: the dense multivariate-normal density of the stacked
sample, every operation in 40-digit mpmath arithmetic (oracle.kalman_numpy.dense_gaussian_loglik(mp_digits=40)).
An algorithm-independent known answer for the one fixture the reference itself value-tests
(tests/test_kalman_filter.py:226-241, there against statsmodels).

    python tests/golden/make_nile_dense.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import kalman_numpy as kn  # noqa: E402
from tests.helpers import nile_inputs  # noqa: E402

out = {}
for n_missing in (0, 5):
    args = nile_inputs(n_missing)
    out[str(n_missing)] = {"loglik": kn.dense_gaussian_loglik(*args, mp_digits=40),
                           "missing_rows": sorted(int(i) for i in __import__("numpy").where(
                               __import__("numpy").isnan(args[0][:, 0, 0]))[0])}
with open(os.path.join(ROOT, "tests", "golden", "nile_dense_loglik.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
