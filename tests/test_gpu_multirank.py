"""Two ranks on two GPUs of one box (NCCL): the multi-GPU resident leg of bench.py (dist.GatherStepGraph) returns, on
every rank, every draw's (logp, grad) row in draw order, identical to a single-GPU evaluation.  Skipped on 1-GPU boxes
(run with `gpurun --gpus 2`); the layout logic itself is covered on CPU by the gloo test in test_cabi_and_host.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pymc_statespace_b200.dist import GatherStepGraph
from pymc_statespace_b200.logp import KalmanLogp
from pymc_statespace_b200.synthetic import arma21_workload
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, n, waves = 4096, 80, 4
spec, y, theta = arma21_workload(B * world, n)
full = KalmanLogp(spec, y, n_draws=B * world, device=dev)
lp, g = full.logp_and_grad(torch.as_tensor(theta, device=dev))
ref = torch.cat([lp[:, None], g], dim=1)
model = KalmanLogp(spec, y, n_draws=B // waves, device=dev)
mine = torch.as_tensor(np.ascontiguousarray(theta[rank * B:(rank + 1) * B]), device=dev)
for use_graph in (False, True):
    print("rank", rank, "use_graph", use_graph, flush=True)
    gsg = GatherStepGraph(model, mine, waves=waves, use_graph=use_graph)
    for _ in range(3):
        gsg()
    torch.cuda.synchronize()
    assert torch.equal(gsg.rows(), ref), (rank, use_graph)
    assert (gsg.graph is not None) == use_graph
dist.barrier()
if rank == 0:
    print("NCCL_GATHER_OK")
dist.destroy_process_group()
"""


def test_gather_step_graph_two_ranks_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker_nccl.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29533", str(script), ROOT],
        capture_output=True, text=True, timeout=180, env=env)
    assert out.returncode == 0 and "NCCL_GATHER_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
