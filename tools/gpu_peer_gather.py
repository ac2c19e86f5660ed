"""torchrun target (N >= 2 GPUs): the peer-memory all-gather of lnL (payne_gather_*) against ncclAllGather on the same
points over many pipelined steps, and the step time of both (dev tool; bench.py uses the same class).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_peer_gather.py"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
from thepayne_b200 import dist as pdist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg, g = load_case(name)
eng = engine_from_config(cfg, precision='parity', device=local)
pg = pdist.PeerGather(eng, B)
thetas = [torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=100 * s + rank))).cuda() for s in range(7)]
# reference: plain NCCL gather of each step
refs = [pdist.gather_equal(eng.lnlike_batch(t)).clone() for t in thetas]
torch.cuda.synchronize()
got = []
for s, t in enumerate(thetas):
    prev = pg.submit(t)
    if prev is not None:
        got.append(prev.clone())
got.append(pg.flush().clone())
torch.cuda.synchronize()
bad = sum(0 if torch.equal(a, b) or torch.allclose(a, b, rtol=0, atol=0, equal_nan=True) else 1 for a, b in zip(got, refs))
print('rank %d: %d steps gathered, %d differ from ncclAllGather, status %d' % (rank, len(got), bad, eng.query('status')), flush=True)
assert bad == 0 and len(got) == len(refs)
# timing: K pipelined steps each way
K = 100
def timed(fn, drain):
    for _ in range(5): fn()
    drain(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    drain(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
th = thetas[0]
t_none = timed(lambda: eng.lnlike_batch(th), lambda: None)
png = pdist.PipelinedGather()
t_nccl = timed(lambda: png.submit(eng.lnlike_batch(th)), lambda: png.flush())
t_peer = timed(lambda: pg.submit(th), lambda: pg.flush())
if rank == 0:
    print('%s B=%d x %d GPUs: %.4f ms/step without gather, %.4f pipelined ncclAllGather, %.4f peer-memory gather' % (
        name, B, world, t_none, t_nccl, t_peer), flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
