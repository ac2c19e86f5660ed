"""Multi-GPU sharding of a live-point batch.

Live points are independent (no cross-point term anywhere in Payne/fitting/likelihood.py), so
the batch is split by contiguous row blocks, one per rank; weights and the observation are
replicated (each rank builds its own context) and the only communication is one all-gather of
the per-point lnL vector -- NCCL over NVLink on GPUs, gloo in the CPU tests of this logic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(B, world, rank):
    """Row block [lo, hi) of ``rank``: sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_lnlike(theta, compute, group=None):
    """Evaluate ``compute(theta_local) -> lnL_local`` on this rank's row block of the replicated
    ``theta`` [B, ndim] and all-gather the full lnL [B] on every rank.

    ``compute`` is ``Engine.lnlike_batch`` in production; the tests inject a CPU stand-in."""
    if not (dist.is_available() and dist.is_initialized()):
        return compute(theta)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B = theta.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    local = compute(theta[lo:hi])
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    nmax = max(h - l for l, h in sizes)
    pad = torch.full((nmax,), float('nan'), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * nmax,), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group) if local.is_cuda else \
        dist.all_gather(list(out.view(world, nmax).unbind(0)), pad, group=group)
    out = out.view(world, nmax)
    return torch.cat([out[r, : h - l] for r, (l, h) in enumerate(sizes)])


def gather_equal(local, group=None):
    """All-gather equally sized per-rank lnL vectors (the weak-scaling bench path)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), local.contiguous(), group=group)
    return out
