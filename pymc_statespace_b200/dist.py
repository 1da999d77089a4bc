"""Multi-GPU plumbing: shard draws over ranks (one process per GPU), gather per-draw logp and gradients.

The path shards with no data-path collective (every draw is an independent recursion);
the only communication is ONE all-gather of ``[B_local, 1 + n_theta]`` float64 per evaluation over
NCCL/NVLink (gloo on CPU for the host-logic tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, rank: int, world: int):
    """Contiguous block of draws owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_logp_grad(logp: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    return torch.cat([logp[:, None], grad], dim=1).contiguous()


def gather_logp_grad(packed_local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank ``[B_local, 1+n_theta]`` blocks into ``[n_total, 1+n_theta]`` (draw order)."""
    if not (dist.is_available() and dist.is_initialized()):
        return packed_local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    width = packed_local.shape[1]
    if all((hi - lo) == (sizes[0][1] - sizes[0][0]) for lo, hi in sizes):
        out = torch.empty((n_total, width), dtype=packed_local.dtype, device=packed_local.device)
        dist.all_gather_into_tensor(out, packed_local, group=group)
        return out
    # ragged shards: pad every block to the largest shard, gather, then drop the padding rows
    biggest = max(hi - lo for lo, hi in sizes)
    padded = torch.zeros((biggest, width), dtype=packed_local.dtype, device=packed_local.device)
    padded[: packed_local.shape[0]] = packed_local
    out = torch.empty((world * biggest, width), dtype=packed_local.dtype, device=packed_local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    out = out.view(world, biggest, width)
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


class GatherStepGraph:
    """Device-resident multi-GPU evaluation: theta[B_local, n_theta] on every rank -> per-draw (logp, grad) rows of ALL
    ranks on every rank.

    The local draws go through ``waves`` sequential waves on one evaluator (``model`` is sized for a single wave).  The
    kernels of a wave (scatter, Lyapunov, forward, adjoint, scatter^T, pack: ~9 launches) are captured once and replayed
    as ONE CUDA graph per wave; the NCCL all-gather of wave k is issued eagerly on a communication stream behind an
    event and overlaps the kernels of wave k + 1, so only the last wave's gather is exposed.
    (The collectives are deliberately NOT captured: a captured NCCL all-gather dead-locked both ranks on this
    torch / NCCL pair - 2-GPU run of round 2.)  Result layout: ``out[wave, rank, i, :]`` holds draw
    ``rank * B_local + wave * h + i`` (h = B_local / waves); ``rows()`` returns it in draw order.

    ``use_graph=False`` launches the kernels eagerly as well.
    """

    def __init__(self, model, theta_dev: torch.Tensor, waves: int = 1, group=None, warmup: int = 3, use_graph: bool = True):
        from .logp import _capture

        self.model, self.theta, self.waves, self.group = model, theta_dev, int(waves), group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        h, nt = model.B, model.spec.n_theta
        if (tuple(theta_dev.shape) != (h * self.waves, nt) or theta_dev.dtype != torch.float64
                or theta_dev.device.type != model.device.type):
            raise ValueError(f"theta_dev: expected a float64 tensor of shape {(h * self.waves, nt)} on {model.device}")
        dev = model.device
        cuda = dev.type == "cuda"  # (a CPU device only occurs in the gloo host-logic tests, with a stub evaluator)
        self._cuda, self._dev = cuda, dev
        self.out = torch.empty((self.waves, self.world, h, 1 + nt), dtype=torch.float64, device=dev)
        self._packed = [torch.empty((h, 1 + nt), dtype=torch.float64, device=dev) for _ in range(self.waves)]
        self.info = torch.zeros((h * self.waves,), dtype=torch.int32, device=dev)
        self._comm = torch.cuda.Stream(device=dev) if cuda else None

        def wave_body(w):
            def run():
                logp, grad = model.logp_and_grad(theta_dev[w * h:(w + 1) * h])
                torch.cat([logp[:, None], grad], dim=1, out=self._packed[w])
                self.info[w * h:(w + 1) * h].copy_(model.info)
                if self.world == 1:
                    self.out[w, 0].copy_(self._packed[w])
            return run

        self._wave = [wave_body(w) for w in range(self.waves)]
        self.graphs = None
        if use_graph and cuda:
            self.graphs = [_capture(dev, run, warmup if w == 0 else 1) for w, run in enumerate(self._wave)]
        for _ in range(max(1, warmup)):
            self()

    @property
    def graph(self):
        """The captured kernel graphs (one per wave) or None when running eagerly."""
        return self.graphs

    def __call__(self) -> torch.Tensor:
        """Enqueue one evaluation on the current stream (no host synchronisation); returns ``out``."""
        h = self.model.B
        cur = torch.cuda.current_stream(self._dev) if self._cuda else None
        if self._cuda and self.world > 1:
            self._comm.wait_stream(cur)          # the previous evaluation's consumers are done with `out`
        for w in range(self.waves):
            if self.graphs is not None:
                self.graphs[w].replay()
            else:
                self._wave[w]()
            if self.world > 1:
                dst = self.out[w].view(self.world * h, self.out.shape[-1])
                if self._cuda:
                    done = cur.record_event()
                    with torch.cuda.stream(self._comm):
                        self._comm.wait_event(done)
                        dist.all_gather_into_tensor(dst, self._packed[w], group=self.group)
                else:
                    dist.all_gather_into_tensor(dst, self._packed[w], group=self.group)
        if self._cuda and self.world > 1:
            cur.wait_stream(self._comm)
        return self.out

    def rows(self) -> torch.Tensor:
        """[world * B_local, 1 + n_theta] in draw order (a copy)."""
        w, r, h, k = self.out.shape
        return self.out.permute(1, 0, 2, 3).reshape(r * w * h, k)
