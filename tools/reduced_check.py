"""dev tool: loglik / gradient of configs[1] (all 65,536 draws) from the reduced ARMA recursion (all four structure promises)
against the general kernels (no promises); the worst draws against a 40-digit dense Gaussian density."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pymc_statespace_b200 import BatchedKalman
from pymc_statespace_b200.logp import KalmanLogp
from pymc_statespace_b200.models import MATRICES
from pymc_statespace_b200.synthetic import arma11_workload
from oracle import kalman_numpy as kn

B, n = 65536, 1000
spec, y, theta = arma11_workload(B, n)
model = KalmanLogp(spec, y, n_draws=B)
mats = model._scatter(torch.as_tensor(theta, device="cuda"))
res = {}
for name, kw in (("reduced", dict(z_unit0=True, h_zero=True, t_companion=True, no_missing=True)), ("general", {})):
    bk = BatchedKalman("standard", n, 2, 1, 1, n_draws=B, **kw)
    out = bk.forward(model.y, *[mats[k] for k in MATRICES], outputs=("loglik",), save_for_backward=True)
    g = bk.backward(wrt=("a0", "P0", "T", "R", "Q"))
    res[name] = (out["loglik"].cpu().numpy(), {k: v.cpu().numpy().reshape(B, -1) for k, v in g.items()})
l1, l0 = res["reduced"][0], res["general"][0]
rel = np.abs(l1 / l0 - 1)
print("loglik: max |rel diff| %.3e at draw %d; 99.9th pct %.3e; max/|max| %.3e" % (rel.max(), rel.argmax(), np.quantile(rel, 0.999), np.abs(l1 - l0).max() / np.abs(l0).max()))
for k in ("a0", "P0", "T", "R", "Q"):
    a, b = res["reduced"][1][k], res["general"][1][k]
    if k == "T":
        a, b = a.reshape(B, 2, 2)[:, :, 0], b.reshape(B, 2, 2)[:, :, 0]
    err = np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-300)
    print("grad %s: max rel %.3e at %d; 99.9th pct %.3e" % (k, err.max(), err.argmax(), np.quantile(err, 0.999)))
# (a 40-digit dense Gaussian density of a worst draw takes ~10 minutes of mpmath at n = 1000: run it on the CPU if needed,
#  oracle.kalman_numpy.dense_gaussian_loglik(..., mp_digits=40))
