"""CPU, world_size 2, gloo: the N>1 path of the likelihood -- row-block sharding of the live
points and the all-gather of the per-point lnL -- with a CPU stand-in for the per-rank compute
(the product has no CPU compute path; on GPUs the same code runs over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from thepayne_b200.dist import gather_equal, shard_bounds, sharded_lnlike
    rng = np.random.default_rng(5)
    theta = torch.from_numpy(rng.standard_normal((B, 7)))
    calls = []

    def compute(t):                       # stand-in for Engine.lnlike_batch
        calls.append(t.shape[0])
        return -0.5 * (t ** 2).sum(1)
    out = sharded_lnlike(theta, compute)
    lo, hi = shard_bounds(B, world, rank)
    ok = torch.equal(out, -0.5 * (theta ** 2).sum(1)) and calls == [hi - lo]
    loc = torch.full((5,), float(rank), dtype=torch.float64)
    g = gather_equal(loc)
    ok = ok and torch.equal(g, torch.arange(world, dtype=torch.float64).repeat_interleave(5))
    # pipelined gather: results come back one submit later, in order, and flush() returns the rest
    from thepayne_b200.dist import PipelinedGather
    pg = PipelinedGather()
    got = []
    for step in range(4):
        r = pg.submit(torch.full((3,), float(10 * step + rank), dtype=torch.float64))
        if r is not None:
            got.append(r.clone())
    got += [r.clone() for r in pg.flush()]
    want = [torch.tensor([10.0 * s + r for r in range(world) for _ in range(3)], dtype=torch.float64) for s in range(4)]
    ok = ok and len(got) == 4 and all(torch.equal(a, b) for a, b in zip(got, want))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [64, 101])
def test_sharded_lnlike_world2(B):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
