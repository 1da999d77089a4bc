#!/usr/bin/env python
"""bench.py - Kalman logp+grad filter-steps/s (draws x T) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--draws B] [--n T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE batched logp+grad evaluation (theta[B,n_theta] -> logp[B], dlogp/dtheta[B,n_theta]) of
BASELINE.json configs[1]: BayesianARMA(1,1), stationary init, 65,536 draws, T=1,000, standard filter.
For N > 1 every rank evaluates its own 65,536 draws (weak scaling) and the per-draw (logp, grad) rows are
all-gathered over NCCL each step.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kalman_logp_grad_filter_steps_per_s"
UNIT = "filter-steps/s"
ALG_FLOPS_PER_STEP = 442.0      # logp+grad, m=2 p=1
TAPE_BYTES_PER_STEP = 40.0      # 8 * (m + m(m+1)/2) written by the forward kernel and read by the adjoint kernel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--draws", type=int, default=65536, help="draws per GPU")
    ap.add_argument("--n", type=int, default=1000, help="time steps T")
    ap.add_argument("--cpu-sample-draws", type=int, default=0, help="draws in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="e2e leg: eager launches instead of the captured CUDA graph")
    return ap.parse_args()


def workload_name(a):
    return f"BayesianARMA(1,1) stationary init, {a.draws} draws/GPU x T={a.n}, standard filter, logp+dlogp/dtheta"


# ---------------------------------------------------------------------------------------------- CPU side
def cpu_inputs(theta, y):
    """Matrix-level inputs of the C oracle port for ARMA(1,1) draws (numpy; restates models/SARIMAX.py update)."""
    import scipy.linalg

    B = theta.shape[0]
    T = np.zeros((B, 2, 2)); T[:, 0, 1] = 1.0; T[:, 0, 0] = theta[:, 3]
    R = np.zeros((B, 2, 1)); R[:, 0, 0] = 1.0; R[:, 1, 0] = theta[:, 4]
    Q = theta[:, 2][:, None, None]
    C = R @ Q @ R.transpose(0, 2, 1)
    # stationary P0: vec(P) = (I - T (x) T)^-1 vec(C)   (same solution as scipy's bilinear method)
    K = np.eye(4)[None] - np.einsum("bij,bkl->bikjl", T, T).reshape(B, 4, 4)
    P0 = np.linalg.solve(K, C.reshape(B, 4, 1)).reshape(B, 2, 2)
    Z = np.array([[1.0, 0.0]]); H = np.zeros((1, 1))
    return y, theta[:, 0:2].copy(), P0, T, Z, H, C


def run_cpu(theta, y, steps, warmup, nthreads=0):
    """Times the plain-C oracle port (forward + adjoint, OpenMP over draws).  Returns (steps/s, cores, ms/step)."""
    from oracle import kalman_c

    args = cpu_inputs(theta, y)
    cores = kalman_c.max_threads() if nthreads == 0 else nthreads
    for _ in range(warmup):
        kalman_c.logp_grad_batch(*args, want_grads=True, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        kalman_c.logp_grad_batch(*args, want_grads=True, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / steps
    return theta.shape[0] * y.shape[0] / dt, cores, dt * 1e3


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from pymc_statespace_b200.synthetic import arma11_workload

    sample = a.cpu_sample_draws or 8192
    _, y, theta = arma11_workload(sample, a.n)
    val, cores, ms = run_cpu(theta, y, a.steps, a.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU arm: each step evaluates a bounded sample of the workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} draws x T={a.n} per step, oracle/kalman_c.c (plain C + OpenMP) - the reference "
                                   "(PyTensor/PyMC) is not installable in this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML in-process, every ~10 ms;
    falls back to one long-running `nvidia-smi -lms 50`)."""

    def __init__(self, index):
        self.index, self.sm, self.smmax, self.reasons = index, [], None, set()
        self.stop = threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run_nvml(self):
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.smmax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop.is_set():
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            self.reasons |= {n for bit, n in names.items() if r & bit}
            self.stop.wait(0.01)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                 "-lms", "50"], stdout=subprocess.PIPE, text=True)
        try:
            while not self.stop.is_set():
                row = [x.strip() for x in proc.stdout.readline().split(",")]
                if len(row) >= 6 and row[0].replace(".", "").isdigit():
                    self.sm.append(float(row[0]))
                    self.smmax = float(row[1])
                    self.reasons |= {n for n, v in zip(names, row[2:6]) if v.lower().startswith("active")}
        finally:
            proc.kill()

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def __enter__(self):
        self.th.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smmax,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------- GPU side
def main():
    a = parse()
    if a.impl == "reference":
        return reference_arm(a)

    import torch
    import torch.distributed as dist

    from pymc_statespace_b200 import _lib, fp64_peak_tflops
    from pymc_statespace_b200.dist import gather_logp_grad, pack_logp_grad
    from pymc_statespace_b200.logp import KalmanLogp
    from pymc_statespace_b200.synthetic import arma11_workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, n = a.draws, a.n
    spec, y, theta_all = arma11_workload(B * world, n)
    theta_h = torch.from_numpy(np.ascontiguousarray(theta_all[rank * B:(rank + 1) * B])).pin_memory()
    theta_d = theta_h.to(dev)
    model = KalmanLogp(spec, y, n_draws=B, filter_type="standard", device=dev)
    n_total = B * world
    out_h = torch.empty((n_total, 1 + spec.n_theta), dtype=torch.float64).pin_memory()

    def step_resident():
        logp, grad = model.logp_and_grad(theta_d)
        packed = pack_logp_grad(logp, grad)
        return gather_logp_grad(packed, n_total) if world > 1 else packed

    # single GPU: the public host-to-host call, captured once as a CUDA graph (KalmanLogp.capture_host_step: H2D of theta
    # from pinned memory, every kernel of the evaluation, D2H of (logp, grad) into pinned memory) and replayed per step;
    # multi-GPU keeps the eager sequence (the all-gather is not captured)
    host_step = None
    if world == 1 and not a.no_graph:
        host_step = model.capture_host_step(theta_h, out_h, chunks=4 if B % 4 == 0 else 1)

    def step_e2e():
        if host_step is not None:
            return host_step()                            # replay + stream synchronise: the numbers are on the host
        th = theta_h.to(dev, non_blocking=True)
        logp, grad = model.logp_and_grad(th)
        packed = pack_logp_grad(logp, grad)
        full = gather_logp_grad(packed, n_total) if world > 1 else packed
        if rank == 0:
            out_h.copy_(full, non_blocking=True)          # the job's result: every draw's (logp, grad) on the host
        else:
            out_h[:B].copy_(packed, non_blocking=True)    # other ranks only read back their own shard
        torch.cuda.current_stream().synchronize()  # the caller needs the numbers on the host
        return out_h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms / steps

    for _ in range(max(a.warmup, 3)):
        step_resident()
    launches0 = lib.kfb_launch_count()
    with ClockSampler(local) as clk:
        ms_step = timed(step_resident, a.steps)
    launches = lib.kfb_launch_count() - launches0
    clocks = clk.summary()
    bad = int((model.info != 0).sum())

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)

    # per-kernel durations of the two recursion kernels, CUDA events on the launching stream
    mats = model._scatter(theta_d)
    from pymc_statespace_b200.models import MATRICES
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    torch.cuda.synchronize()
    for e in ev:
        e[0].record()
        model.kalman.forward(model.y, *[mats[k] for k in MATRICES], outputs=("loglik",), save_for_backward=True)
        e[1].record()
        model.kalman.backward(wrt=("a0", "P0", "T", "R", "Q"))
        e[2].record()
    torch.cuda.synchronize()
    ms_fwd = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))   # rqr + forward kernel
    ms_bwd = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))   # adjoint kernel + rqr adjoint

    steps_per_eval = B * n
    value = steps_per_eval * world / (ms_step * 1e-3)
    e2e_value = steps_per_eval * world / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        fp64_peak = fp64_peak_tflops(dev)
        fp64_peak_distinct = fp64_peak_tflops(dev, distinct_operands=True)
        traffic = {}
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this exact workload
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            if B == 65536 and n == 1000:
                traffic = {k: tj[k]["dram_bytes_read"] + tj[k]["dram_bytes_write"] for k in ("adjoint", "forward")}
        except Exception:
            pass
        tape_bytes = B * (n - 1) * TAPE_BYTES_PER_STEP
        bwd_gbs = tape_bytes / (ms_bwd * 1e-3) / 1e9
        fwd_gbs = tape_bytes / (ms_fwd * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "k_states": 2, "k_endog": 1, "n_theta": spec.n_theta,
                       "gradient": "theta-level [B,5] (scatter + Lyapunov + Kalman adjoint)",
                       "l2": "no flush: each step streams a %.2f GB tape (write in forward, read in adjoint) >> 126 MB L2"
                             % (tape_bytes / 1e9),
                       "parallelism": f"draws sharded x{world}, all-gather of [B,{1 + spec.n_theta}] f64" if world > 1 else "1 GPU",
                       "draws_with_info": bad},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(theta_h.numel() * 8 * world),
                    "d2h_bytes_per_step": int(out_h.numel() * 8 + (world - 1) * B * (1 + spec.n_theta) * 8),
                    "path": ("KalmanLogp.capture_host_step: pinned H2D + evaluation + pinned D2H replayed as one CUDA graph "
                             "(4 parallel draw-chunk branches: copies overlap kernels), stream-synchronised every step") if host_step is not None else
                            "eager: pinned H2D, logp_and_grad, all-gather, pinned D2H, stream-synchronised every step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "kf_thread_kernel<2,1,MK_STD,2> (adjoint recursion; the longest kernel)",
                         "achieved": bwd_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": bwd_gbs / hbm_peak,
                         "peak_source": peak_src, "traffic": traffic.get("adjoint"), "ms_per_launch": ms_bwd,
                         "algorithmic_bytes_per_launch": tape_bytes,
                         "note": "algorithmic bytes = 40 B/step tape read; ms_per_launch includes the ~8 us R Q R^T adjoint "
                                 "helper launched with it; this kernel is fp64-latency bound, see the fp64 block"},
            "roofline_forward": {"bound": "hbm", "kernel": "kf_thread_kernel<2,1,MK_STD,0> (forward: loglik + tape)",
                                 "achieved": fwd_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fwd_gbs / hbm_peak,
                                 "traffic": traffic.get("forward"), "ms_per_launch": ms_fwd,
                                 "algorithmic_bytes_per_launch": tape_bytes},
            "fp64": {"peak_tflops_measured": fp64_peak, "peak_tflops_distinct_operands": fp64_peak_distinct, "achieved_tflops": ALG_FLOPS_PER_STEP * steps_per_eval /
                     ((ms_fwd + ms_bwd) * 1e-3) / 1e12,
                     "frac": ALG_FLOPS_PER_STEP * steps_per_eval / ((ms_fwd + ms_bwd) * 1e-3) / 1e12 / fp64_peak,
                     "flops_per_step": ALG_FLOPS_PER_STEP},
        }
        if world == 1 and not a.no_cpu_baseline:
            from pymc_statespace_b200.synthetic import arma11_workload as wl

            sample = a.cpu_sample_draws or 16384
            _, ys, ths = wl(sample, n)
            reps = 3
            v, cores, ms = run_cpu(ths, ys, reps, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} draws x T={n}, {reps} timed repeats ({ms:.0f} ms each), "
                                              "oracle/kalman_c.c forward+adjoint, OpenMP over draws"}
            # context only: the per-step Python restatement (closest analogue of the reference's PyTensor scan VM,
            # which BASELINE.md estimates at 7e3-3e4 steps/s/core) on one draw, forward + torch-autograd backward
            try:
                from oracle import models as om

                t0 = time.perf_counter()
                om.logp_and_grad_theta(lambda t: om.arma_matrices(t, (1, 1), True), ths[0], ys[:, :, None])
                line["cpu_baseline"]["python_oracle_steps_per_s_1core"] = n / (time.perf_counter() - t0)
            except Exception:
                pass
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
