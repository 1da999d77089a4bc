"""Developer probe: end-to-end step (pinned H2D -> logp+grad -> pinned D2H) eager vs replayed as ONE CUDA graph."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pymc_statespace_b200.logp import KalmanLogp
from pymc_statespace_b200.synthetic import arma11_workload
from pymc_statespace_b200.dist import pack_logp_grad

B, n = 65536, 1000
spec, y, theta = arma11_workload(B, n)
dev = torch.device("cuda:0")
theta_h = torch.from_numpy(np.ascontiguousarray(theta)).pin_memory()
out_h = torch.empty((B, 1 + spec.n_theta), dtype=torch.float64).pin_memory()
model = KalmanLogp(spec, y, n_draws=B, filter_type="standard", device=dev)

def eager():
    th = theta_h.to(dev, non_blocking=True)
    logp, grad = model.logp_and_grad(th)
    out_h.copy_(pack_logp_grad(logp, grad), non_blocking=True)
    torch.cuda.current_stream().synchronize()

def timeit(fn, k=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3

ms_eager = timeit(eager)
ref = out_h.clone()
static_th = torch.empty((B, spec.n_theta), dtype=torch.float64, device=dev)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        static_th.copy_(theta_h, non_blocking=True)
        lp, g = model.logp_and_grad(static_th)
        out_h.copy_(pack_logp_grad(lp, g), non_blocking=True)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    static_th.copy_(theta_h, non_blocking=True)
    lp, g = model.logp_and_grad(static_th)
    out_h.copy_(pack_logp_grad(lp, g), non_blocking=True)
out_h.zero_()
def replay():
    graph.replay()
    torch.cuda.current_stream().synchronize()
ms_graph = timeit(replay)
ok = bool(torch.equal(out_h, ref))
print(json.dumps({"ms_eager": ms_eager, "ms_graph": ms_graph, "same_result": ok}))
