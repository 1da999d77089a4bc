"""dev tool: full-output forward pass (row f2) timing.  usage: python tools/full_bench.py [m p r n draws]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from pymc_statespace_b200 import _lib
if os.environ.get("KFB_LIB"): _lib.LIB_PATH = os.environ["KFB_LIB"]
from pymc_statespace_b200 import BatchedKalman
from tests.helpers import random_system

m, p, r, n, B = (int(a) for a in sys.argv[1:6]) if len(sys.argv) > 5 else (2, 1, 1, 1000, 65536)
ALL = ("filtered_states", "predicted_states", "filtered_covs", "predicted_covs", "loglik", "ll_obs")
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")  # noqa: E731
rng = np.random.default_rng(0)
y, a0, P0, T, Z, R, H, Q = random_system(rng, m, p, r, n)
Tb = np.repeat(T[None], B, 0) * (1 + 0.01 * rng.normal(size=(B, 1, 1)))
args = (dev(y[..., 0]), dev(a0[:, 0]), dev(P0), dev(Tb), dev(Z), dev(R), dev(H), dev(Q))
bk = BatchedKalman("standard", n, m, p, r, n_draws=B)
U = B
bufs = {"filtered_states": (U, n, m), "predicted_states": (U, n + 1, m), "filtered_covs": (U, n, m, m),
        "predicted_covs": (U, n + 1, m, m), "loglik": (U,), "ll_obs": (U, n), "info": (U,)}
out = {k: torch.empty(s, dtype=torch.int32 if k == "info" else torch.float64, device="cuda") for k, s in bufs.items()}
best = 1e30
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bk.forward(*args, outputs=ALL, out=out); e1.record(); torch.cuda.synchronize()
    if it: best = min(best, e0.elapsed_time(e1))
bytes_step = 8 * (2 * m + 2 * m * m) + 8
gb = B * n * bytes_step / 1e9
print(json.dumps({"m": m, "p": p, "draws": B, "n": n, "ms": best, "GBps": gb / (best * 1e-3), "frac_of_6543.7": gb / (best * 1e-3) / 6543.7,
                  "steps_per_s": B * n / (best * 1e-3), "bad": int((out["info"] != 0).sum())}))
if os.environ.get("KFB_SMOOTHER"):
    from pymc_statespace_b200.engine import rts_smoother
    best = 1e30
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ss, sc = rts_smoother(args[3], args[5], args[7], out["filtered_states"], out["filtered_covs"]); e1.record()
        torch.cuda.synchronize()
        if it: best = min(best, e0.elapsed_time(e1))
    bytes_s = 8 * 2 * (m + m * m)  # read filtered, write smoothed moments
    print(json.dumps({"smoother_ms": best, "steps_per_s": B * n / (best * 1e-3), "GBps": B * n * bytes_s / 1e9 / (best * 1e-3)}))
