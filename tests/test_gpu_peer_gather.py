"""The library's own all-gather of lnL over peer memory (payne_gather_*, thepayne_b200.dist.PeerGather) against the plain
likelihood call: rotation of the three buffers, the one-step-late hand-over, flush.  On a one-GPU box the group has one
rank (same kernels and flags, no peer mapping); with two or more GPUs a second test spawns one process per GPU and checks
every rank's gathered vectors against ncclAllGather (tools/gpu_peer_gather.py is the same check as a torchrun target; it was
run on 2 and 8 B200s)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import load_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_peer_gather_single_rank_group():
    import torch.distributed as dist
    from thepayne_b200 import dist as pdist
    from thepayne_b200.engine import engine_from_config
    dist.init_process_group('nccl', init_method='tcp://127.0.0.1:%d' % _free_port(), world_size=1, rank=0,
                            device_id=torch.device('cuda', 0))
    try:
        cfg, g = load_case('mini_joint')
        eng = engine_from_config(cfg, precision='parity')
        B = 37
        pg = pdist.PeerGather(eng, B)
        thetas = [torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=s))).cuda() for s in range(5)]
        ref = [eng.lnlike_batch(t).clone() for t in thetas]
        got = []
        for t in thetas:
            prev = pg.submit(t)
            if prev is not None:
                got.append(prev.clone())
        got.append(pg.flush().clone())
        torch.cuda.synchronize()
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
        assert eng.query('status') == 0
        with pytest.raises(ValueError):
            pg.submit(thetas[0][:5])
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_peer_gather_two_ranks_against_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr',
           '127.0.0.1', '--master-port', str(_free_port()), os.path.join(ROOT, 'tools', 'gpu_peer_gather.py'), 'mini_spec', '512']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('0 differ from ncclAllGather') == n
