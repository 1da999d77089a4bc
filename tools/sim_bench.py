"""dev tool: timings of the simulation entry points (row f4) and of the Lyapunov solver at larger k_states."""
import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from pymc_statespace_b200.simulation import simulate_statespace, conditional_simulation
from pymc_statespace_b200.engine import lyapunov_forward
from tests.helpers import random_system

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")
out = []
for (m, p, r, n, B, S) in ((2, 1, 1, 1000, 4096, 16), (6, 3, 3, 1000, 1024, 16), (30, 1, 3, 500, 256, 4)):
    rng = np.random.default_rng(m)
    y, a0, P0, T, Z, R, H, Q = random_system(rng, m, p, r, n, scale_T=0.2)
    Tb = np.repeat(T[None], B, 0)
    zs = torch.randn((B * S, n, r), dtype=torch.float64, device="cuda"); zo = torch.randn((B * S, n, p), dtype=torch.float64, device="cuda")
    ms = timed(lambda: simulate_statespace(dev(Tb), dev(Z), dev(R), dev(H), dev(Q), n, n_simulations=S, z_state=zs, z_obs=zo))
    A = rng.normal(size=(B * S, n, m, m)) if m <= 6 else None
    ms2 = None
    if A is not None:
        covs = dev(A @ np.swapaxes(A, -1, -2) + np.eye(m)); mus = dev(rng.normal(size=(B * S, n, m)))
        z = torch.randn((B * S, n, m), dtype=torch.float64, device="cuda"); jit = torch.full((B * S,), 1e-10, dtype=torch.float64, device="cuda")
        ms2 = timed(lambda: conditional_simulation(mus, covs, 1, z=z, jitter=jit))
    Bl = 65536 if m <= 6 else 8192
    ms3 = timed(lambda: lyapunov_forward(dev(np.repeat(T[None], Bl, 0)), dev(R), dev(Q)))
    out.append({"m": m, "simulate_ms": ms, "traj_steps_per_s": B * S * n / (ms * 1e-3), "mvn_draws_ms": ms2,
                "mvn_steps_per_s": None if ms2 is None else B * S * n / (ms2 * 1e-3), "lyapunov_ms": ms3, "lyapunov_draws": Bl})
print(json.dumps(out, indent=1))
