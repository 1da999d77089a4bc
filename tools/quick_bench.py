"""Quick device timing of forward+backward at matrix level (dev tool; bench.py is the contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import os
from pymc_statespace_b200 import _lib
if os.environ.get('KFB_LIB'): _lib.LIB_PATH=os.environ['KFB_LIB']
from pymc_statespace_b200 import BatchedKalman, fp64_peak_tflops

def arma11(B, n, seed=1):
    rng = np.random.default_rng(seed)
    rho = rng.uniform(-0.95, 0.95, B); th = rng.normal(0, 0.5, B); sig = np.exp(rng.normal(0, 0.3, B))
    T = np.zeros((B, 2, 2)); T[:, 0, 0] = rho; T[:, 0, 1] = 1.0
    R = np.zeros((B, 2, 1)); R[:, 0, 0] = 1; R[:, 1, 0] = th
    Q = sig[:, None, None].copy(); Z = np.array([[1.0, 0.0]]); H = np.zeros((1, 1))
    a0 = rng.normal(size=(B, 2))
    P0 = np.tile(np.eye(2), (B, 1, 1)) * 2.0
    rng0 = np.random.default_rng(0); y = np.zeros(n); e = rng0.normal(size=n + 1)
    for t in range(n): y[t] = 0.6 * (y[t - 1] if t else 0) + e[t + 1] + 0.3 * e[t]
    return y[:, None], a0, P0, T, Z, R, H, Q

WRT = tuple(os.environ.get('KFB_WRT', 'a0,P0,T,R,Q').split(','))  # default: what the ARMA theta-level path asks for

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    dev = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device="cuda")
    y, a0, P0, T, Z, R, H, Q = map(dev, arma11(B, n))
    print("fp64 peak TFLOP/s:", fp64_peak_tflops())
    for force, gen in ((False, False), (False, True)) + (((True, False),) if B <= 16384 else ()):
        st = os.environ.get('KFB_STRUCT', '1') == '1' and not gen and not force   # ARMA: Z = [1, 0], H = 0
        bk = BatchedKalman("standard", n, 2, 1, 1, n_draws=B, force_coop=force, generic_adjoint=gen, z_unit0=st, h_zero=st, t_companion=st, no_missing=st)
        # the CPU must run AHEAD of the GPU or the events also measure launch latency: enqueue several evaluations
        # back to back, time the later ones
        for it in range(3):
            out = bk.forward(y, a0, P0, T, Z, R, H, Q, outputs=("loglik",), save_for_backward=True)
            g = bk.backward(wrt=WRT)
        torch.cuda.synchronize()
        evs = []
        for it in range(6):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out = bk.forward(y, a0, P0, T, Z, R, H, Q, outputs=("loglik",), save_for_backward=True)
            e[1].record()
            g = bk.backward(wrt=WRT)
            e[2].record()
            evs.append(e)
        torch.cuda.synchronize()
        tf = min(e[0].elapsed_time(e[1]) for e in evs[2:])
        tb = min(e[1].elapsed_time(e[2]) for e in evs[2:])
        evs = []
        for it in range(8):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record()
            bk.forward(y, a0, P0, T, Z, R, H, Q, outputs=("loglik",), save_for_backward=False)
            e[1].record()
            evs.append(e)
        torch.cuda.synchronize()
        e = min(evs[3:], key=lambda ev: ev[0].elapsed_time(ev[1]))
        print(f"   forward without tape: {e[0].elapsed_time(e[1]):.3f} ms")
        print(f"coop={force} generic_adjoint={gen} struct={st} B={B} n={n} fwd {tf:.3f} ms bwd {tb:.3f} ms  -> {B*n/((tf+tb)*1e-3):.3e} steps/s; "
              f"ll[0]={float(out['loglik'][0]):.6f} gT[0]={g['T'][0].flatten().tolist()} info!=0: {int((out['info']!=0).sum())}")
        if force and B > 16384: break
main()
