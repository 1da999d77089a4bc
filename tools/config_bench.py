"""Supplementary timings of the other BASELINE.json configs (dev tool; bench.py is the contract).
usage: python tools/config_bench.py varmax|seasonal|arma21 [draws] [n] [filter] [corrected]"""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import os
from pymc_statespace_b200 import _lib
if os.environ.get("KFB_LIB"): _lib.LIB_PATH = os.environ["KFB_LIB"]
from pymc_statespace_b200.logp import KalmanLogp
from pymc_statespace_b200 import synthetic as syn

FLOPS = {("varmax", "standard"): 9234, ("varmax", "cholesky"): 9234, ("varmax", "univariate"): 5112,
         ("seasonal", "standard"): 684202, ("seasonal", "steady_state"): 684202, ("arma21", "standard"): 442}

def main():
    cfg = sys.argv[1]
    draws = int(sys.argv[2]) if len(sys.argv) > 2 else None
    n = int(sys.argv[3]) if len(sys.argv) > 3 else None
    kind = sys.argv[4] if len(sys.argv) > 4 else "standard"
    strict = not (len(sys.argv) > 5 and sys.argv[5] == "corrected")
    if cfg == "varmax":
        spec, y, theta = syn.varmax20_workload(draws or 16384, n or 1000)
    elif cfg.startswith("seasonal"):  # "seasonal" = period 29 (config 4); "seasonal12" = period 12 (13 states), ...
        spec, y, theta = syn.trend_seasonal_workload(draws or 1024, n or 2000, period=int(cfg[8:] or 29))
    else:
        spec, y, theta = syn.arma21_workload(draws or (1 << 20), n or 1000)
    B, T = theta.shape[0], y.shape[0]
    model = KalmanLogp(spec, y, n_draws=B, filter_type=kind, strict_reference=strict,
                       pad_to_fused=not os.environ.get("KFB_NO_PAD"))
    th = torch.as_tensor(theta, device="cuda")
    times = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); logp, grad = model.logp_and_grad(th); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times[1:])
    sps = B * T / (ms * 1e-3)
    fl = FLOPS.get((cfg, kind), 0)
    print(json.dumps({"config": cfg, "filter": kind, "strict": strict, "draws": B, "n": T, "m": spec.k_states, "p": spec.k_endog,
                      "ms": round(ms, 3), "steps_per_s": sps, "alg_tflops": sps * fl / 1e12, "bad": int((model.info != 0).sum()),
                      "finite": bool(torch.isfinite(logp).all()), "logp0": float(logp[0])}))
main()
