"""Generates tests/golden/oracle_golden.npz from the CPU oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference itself cannot be imported in this image
(PyTensor/PyMC absent), so these vectors freeze the ORACLE, not the reference."""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from oracle import kalman_numpy as kn  # noqa: E402
from oracle import kalman_torch as kt  # noqa: E402
from tests.helpers import random_system  # noqa: E402

out = {}
for kind, seed, m, p, r, n, miss in [("standard", 1, 2, 1, 1, 30, 3), ("standard", 2, 6, 3, 3, 20, 2),
                                     ("univariate", 3, 6, 3, 3, 20, 2), ("cholesky", 4, 4, 1, 2, 25, 0),
                                     ("single", 5, 3, 1, 1, 25, 4), ("steady_state", 6, 3, 2, 2, 20, 0)]:
    args = random_system(np.random.default_rng(seed), m, p, r, n, n_missing=miss)
    o = kn.kalman_filter(kind, *args)
    _, g = kt.loglik_and_grads(kind, *args)
    key = f"{kind}-{seed}-{m}-{p}-{r}-{n}-{miss}"
    out[key + "_ll"], out[key + "_fs"], out[key + "_gT"] = o[4], o[0], g["T"]
np.savez(os.path.join("tests", "golden", "oracle_golden.npz"), **out)
print("wrote", len(out), "arrays")
