"""Developer probe: host-to-host step as one CUDA graph with C parallel branches (draw chunks), so that each chunk's
H2D / D2H copies overlap the other chunks' kernels."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pymc_statespace_b200.logp import KalmanLogp
from pymc_statespace_b200.synthetic import arma11_workload

B, n = 65536, 1000
spec, y, theta = arma11_workload(B, n)
dev = torch.device("cuda:0")
theta_h = torch.from_numpy(np.ascontiguousarray(theta)).pin_memory()

def build(C):
    h = B // C
    out_h = torch.empty((B, 1 + spec.n_theta), dtype=torch.float64).pin_memory()
    subs = [KalmanLogp(spec, y, n_draws=h, filter_type="standard", device=dev) for _ in range(C)]
    th_d = [torch.empty((h, spec.n_theta), dtype=torch.float64, device=dev) for _ in range(C)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(C)]
    def body():
        cur = torch.cuda.current_stream()
        for c in range(C):
            s = streams[c]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                th_d[c].copy_(theta_h[c * h:(c + 1) * h], non_blocking=True)
                lp, g = subs[c].logp_and_grad(th_d[c])
                out_h[c * h:(c + 1) * h].copy_(torch.cat([lp[:, None], g], dim=1), non_blocking=True)
        for c in range(C):
            cur.wait_stream(streams[c])
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    def replay():
        graph.replay(); torch.cuda.current_stream().synchronize()
    return replay, out_h, (subs, th_d, streams, graph)

def timeit(fn, k=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3

res, ref = {}, None
for C in (1, 2, 4, 8, 16):
    rp, out_h, keep = build(C)
    res[C] = timeit(rp)
    if ref is None: ref = out_h.clone()
    else: res[f"same_{C}"] = bool(torch.allclose(out_h, ref, rtol=1e-12, atol=0))
print(json.dumps(res))
