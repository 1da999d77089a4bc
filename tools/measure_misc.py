"""dev tool: numbers that existed as code but had never been timed.
  (a) full-output forward (the reference's six outputs per step) at k_states 2 / 6 / 30: GB/s written vs the HBM copy peak;
  (b) the plugin seam at B = 1 and B = 4 (PyMC's default 4 chains): microseconds per logp+grad through
      filters.StandardFilter / torch_op.kalman_logp_grads with numpy inputs, T = 100 / 1000, next to the C port on 1 core;
  (c) time-varying T at k_states = 2 (generic run-time-dims kernels) vs the static hot path.
    python tools/measure_misc.py > gpurun_out/measure_misc.json"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymc_statespace_b200 import BatchedKalman  # noqa: E402
from pymc_statespace_b200.filters import FILTER_FACTORY  # noqa: E402
from pymc_statespace_b200.torch_op import kalman_logp_grads  # noqa: E402
from tests.helpers import random_system  # noqa: E402

HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6543.7) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6543.7
dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")  # noqa: E731
ALL = ("filtered_states", "predicted_states", "filtered_covs", "predicted_covs", "loglik", "ll_obs")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def system(m, p, r, n, B, seed=0):
    rng = np.random.default_rng(seed)
    y, a0, P0, T, Z, R, H, Q = random_system(rng, m, p, r, n)
    Tb = np.repeat(T[None], B, 0) * (1 + 0.01 * rng.normal(size=(B, 1, 1)))
    return dev(y[..., 0]), dev(a0[:, 0]), dev(P0), dev(Tb), dev(Z), dev(R), dev(H), dev(Q)


out = {"hbm_peak_gbs": HBM}
# ---------------------------------------------------------------- (a) full-output forward
full = []
for (m, p, r, n, B) in ((2, 1, 1, 1000, 65536), (6, 3, 3, 1000, 16384), (30, 1, 3, 500, 1024)):
    args = system(m, p, r, n, B)
    bk = BatchedKalman("standard", n, m, p, r, n_draws=B)
    ms = timed(lambda: bk.forward(*args, outputs=ALL), 3)
    bytes_step = 8 * (2 * m + 2 * m * m) + 8
    gb = B * n * bytes_step / 1e9
    full.append({"k_states": m, "k_endog": p, "draws": B, "n": n, "ms": ms, "bytes_per_step": bytes_step, "GB_written": gb,
                 "GBps": gb / (ms * 1e-3), "frac_of_hbm_peak": gb / (ms * 1e-3) / HBM, "steps_per_s": B * n / (ms * 1e-3),
                 "note": "includes torch.empty of the six output tensors"})
    del bk, args
    torch.cuda.empty_cache()
out["full_output_forward"] = full
# ---------------------------------------------------------------- (b) plugin seam
seam = []
from oracle import kalman_c  # noqa: E402  (dev tool: the CPU port as the side-by-side number)
for n in (100, 1000):
    rng = np.random.default_rng(1)
    y, a0, P0, T, Z, R, H, Q = random_system(rng, 2, 1, 1, n)
    flt = FILTER_FACTORY["standard"]()
    reps = 20
    flt.build_graph(y, a0, P0, T, Z, R, H, Q)               # first call of a geometry captures its CUDA graph
    t0 = time.perf_counter()
    for _ in range(reps):
        flt.build_graph(y, a0, P0, T, Z, R, H, Q)           # numpy in -> six numpy outputs (what the Op's perform does)
    us_fwd = (time.perf_counter() - t0) / reps * 1e6
    ts = [dev(v) for v in (y, a0, P0, T, Z, R, H, Q)]
    names = ("a0", "P0", "T", "Z", "R", "H", "Q")
    kalman_logp_grads(flt, ts[0], dict(zip(names, ts[1:])))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        ll, g = kalman_logp_grads(flt, ts[0], dict(zip(names, ts[1:])))
        float(ll)                                            # the sampler needs the number on the host
    us_grad_eager = (time.perf_counter() - t0) / reps * 1e6
    # what KalmanFilterGradOp.perform does: numpy in, numpy out, one replayed CUDA graph (seam.SeamGraph)
    from pymc_statespace_b200.seam import logp_grads_numpy
    arrays = dict(zip(("data",) + names, (y, a0, P0, T, Z, R, H, Q)))
    logp_grads_numpy(flt, arrays)
    t0 = time.perf_counter()
    for _ in range(5 * reps):
        logp_grads_numpy(flt, arrays)
    us_grad = (time.perf_counter() - t0) / (5 * reps) * 1e6
    # B = 4 chains batched in one call
    bk4 = BatchedKalman("standard", n, 2, 1, 1, n_draws=4)
    rep4 = lambda x: x[None].repeat(4, *([1] * x.ndim)).contiguous()  # noqa: E731
    a4 = [ts[0][..., 0]] + [rep4(ts[1][:, 0]), rep4(ts[2]), rep4(ts[3]), ts[4], rep4(ts[5]), ts[6], rep4(ts[7])]
    def both():
        o = bk4.forward(*a4, outputs=("loglik",), save_for_backward=True); bk4.backward(); return o
    both(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        o = both(); o["loglik"].cpu()
    us_b4 = (time.perf_counter() - t0) / reps * 1e6
    C = R @ Q @ R.T
    t0 = time.perf_counter()
    for _ in range(reps):
        kalman_c.logp_grad_batch(y[..., 0], a0.reshape(1, 2), P0[None], T[None], Z, H, C[None], nthreads=1)
    us_c = (time.perf_counter() - t0) / reps * 1e6
    seam.append({"n": n, "us_forward_six_outputs_numpy_in_out": us_fwd, "us_logp_grad_B1": us_grad, "us_logp_grad_B1_eager_torch": us_grad_eager, "us_logp_grad_B4_batched": us_b4,
                 "us_logp_grad_c_port_1core_B1": us_c})
out["plugin_seam_k_states_2"] = seam
# ---------------------------------------------------------------- (c) time-varying T
B, n = 16384, 1000
args = list(system(2, 1, 1, n, B))
bk = BatchedKalman("standard", n, 2, 1, 1, n_draws=B)
def static():
    bk.forward(*args, outputs=("loglik",), save_for_backward=True); bk.backward(wrt=("a0", "P0", "T", "R", "Q"))
ms_static = timed(static, 3)
Ttv = args[3][:, None].repeat(1, n, 1, 1).contiguous()
bktv = BatchedKalman("standard", n, 2, 1, 1, n_draws=B, time_varying=("T",))
atv = list(args); atv[3] = Ttv
def tv():
    bktv.forward(*atv, outputs=("loglik",), save_for_backward=True); bktv.backward(wrt=("a0", "P0", "T", "R", "Q"))
ms_tv = timed(tv, 2)
out["time_varying_T_k_states_2"] = {"draws": B, "n": n, "ms_static_fwd_bwd": ms_static, "ms_time_varying_fwd_bwd": ms_tv,
                                    "slowdown": ms_tv / ms_static, "steps_per_s_time_varying": B * n / (ms_tv * 1e-3)}
print(json.dumps(out, indent=1))
