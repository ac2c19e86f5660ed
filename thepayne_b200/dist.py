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


class PipelinedGather(object):
    """All-gather of equally sized per-rank lnL vectors that overlaps with the NEXT batch's kernels.

    The gather of a 32 KB vector is latency (~30 us of NCCL launch + ring), not bandwidth; issued on
    the compute stream it sits on the critical path of every step.  Here it runs on a side stream: the
    caller submits the local vector, immediately goes on to the next batch, and collects the gathered
    vector one step later (two rotating output buffers).  ``depth`` results are kept in flight.

        pg = PipelinedGather()
        for theta in batches:
            done = pg.submit(engine.lnlike_batch(theta))   # gathered lnL of the PREVIOUS batch, or None
        last = pg.flush()
    """

    def __init__(self, group=None, depth=2):
        self.group, self.depth = group, int(depth)
        self.bufs, self.events, self.pending = [None] * self.depth, [None] * self.depth, []
        self.i = 0
        self.comm = None

    def submit(self, local):
        if not (dist.is_available() and dist.is_initialized()):
            prev = self.pending.pop(0) if self.pending else None
            self.pending.append(local)
            return prev
        world = dist.get_world_size(self.group)
        k = self.i % self.depth
        self.i += 1
        if self.bufs[k] is None or self.bufs[k].shape[0] != world * local.shape[0]:
            self.bufs[k] = torch.empty((world * local.shape[0],), dtype=local.dtype, device=local.device)
        out = self.bufs[k]
        if local.is_cuda:
            if self.comm is None:
                self.comm = torch.cuda.Stream(device=local.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(local.device))
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ready)
                dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
                done = torch.cuda.Event()
                done.record(self.comm)
            local.record_stream(self.comm)
            self.events[k] = done
        else:
            dist.all_gather(list(out.view(world, -1).unbind(0)), local.contiguous(), group=self.group)
            self.events[k] = None
        self.pending.append(k)
        if len(self.pending) >= self.depth:
            return self._collect(self.pending.pop(0))
        return None

    def _collect(self, k):
        if not isinstance(k, int):
            return k
        if self.events[k] is not None:
            torch.cuda.current_stream(self.bufs[k].device).wait_event(self.events[k])
        return self.bufs[k]

    def flush(self):
        """Gathered vectors still in flight, oldest first (the compute stream waits for each)."""
        out = [self._collect(k) for k in self.pending]
        self.pending = []
        return out


class _DevArray(object):
    """A raw device pointer dressed up for ``torch.as_tensor`` (no copy, no ownership)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {'shape': (int(n),), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 3}


class PeerGather(object):
    """All-gather of the per-rank lnL vectors over NVLink peer memory, done by the likelihood library itself
    (``payne_gather_*``, include/payne_b200.h) instead of a collective-library call behind every step.

    Every rank's context holds three rotating buffers of ``world * slots`` doubles whose CUDA IPC handles are
    exchanged once (here through ``all_gather_object``); after that ``submit(theta)`` is one stream-ordered call:
    the tail kernel writes this rank's lnL into its slice, and one small kernel stores the slice into every peer's
    buffer through the peer mappings and raises a flag there.  No kernel of another library has to find room beside
    the persistent tail kernel, which is what kept the side-stream ncclAllGather of ``PipelinedGather`` on the critical
    path (8 GPUs: +29 us per 0.85 ms step).

        pg = PeerGather(engine, slots=B_local)
        for theta in batches:                       # every rank, same number of calls
            prev = pg.submit(theta)                 # gathered lnL [world * B_local] of the PREVIOUS batch (None first)
            ...consume prev on the current stream before the next submit...
        last = pg.flush()

    Raises ``RuntimeError`` on EVERY rank if any rank cannot map its peers (no P2P path); callers fall back to NCCL."""

    def __init__(self, engine, slots, group=None):
        import ctypes
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('PeerGather needs an initialised process group')
        self.engine, self.group, self.slots = engine, group, int(slots)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device('cuda', engine.device)
        buf = ctypes.create_string_buffer(_lib.GATHER_HANDLE_BYTES)
        rc = engine.lib.payne_gather_create(engine._ctx, self.world, self.rank, self.slots, buf)
        handles = [None] * self.world
        dist.all_gather_object(handles, buf.raw if rc == 0 else None, group=group)
        ok = rc == 0 and all(h is not None for h in handles)
        if ok:
            allh = b''.join(handles)
            ok = engine.lib.payne_gather_connect(engine._ctx, allh) == 0
        flag = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) != 1:
            raise RuntimeError('peer-memory gather unavailable on at least one rank: ' +
                               (engine.lib.payne_last_error() or b'').decode())
        self.n = self.world * self.slots

    def _wrap(self, ptr):
        return torch.as_tensor(_DevArray(ptr, self.n), device=self.device) if ptr else None

    def submit(self, theta):
        """theta: this rank's [slots, ndim] float64 CUDA tensor.  Returns the gathered lnL of the previous submit."""
        import ctypes
        from . import _lib
        th = self.engine._theta_dev(theta)
        if th.shape[0] != self.slots:
            raise ValueError('every submit carries exactly %d points per rank' % self.slots)
        prev = ctypes.c_void_p()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.engine.lib.payne_lnlike_batch_gather(self.engine._ctx, th.data_ptr(), th.shape[0], th.shape[1],
                                                             st, ctypes.byref(prev)))
        return self._wrap(prev.value)

    def flush(self):
        """The gathered lnL of the last submit (the current stream waits for every rank's slice)."""
        import ctypes
        from . import _lib
        last = ctypes.c_void_p()
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.engine.lib.payne_gather_flush(self.engine._ctx, st, ctypes.byref(last)))
        return self._wrap(last.value)
