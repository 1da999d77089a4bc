"""Tiny k_states=6 theta-level logp+grad (rows kernels) for compute-sanitizer runs:
   compute-sanitizer --tool racecheck python tools/rows_racecheck.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymc_statespace_b200.synthetic import varmax20_workload
from pymc_statespace_b200.logp import KalmanLogp

spec, y, theta = varmax20_workload(n_draws=48, n=24, seed=1)
m = KalmanLogp(spec, y, n_draws=48, filter_type="standard")
th = torch.as_tensor(theta, device="cuda")
lp, g = m.logp_and_grad(th)
torch.cuda.synchronize()
print("logp sum", float(lp.sum()), "grad abs sum", float(g.abs().sum()))
