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
