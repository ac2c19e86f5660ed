"""B=1..64 latency split into emulator / tail (event timing inside the library) -- dev tool."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
for B in [1, 64, 444]:
    tht = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=B))).cuda()
    eng.set('timing', 0)
    for _ in range(5): eng.lnlike_batch(tht)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): eng.lnlike_batch(tht)
    torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 50
    eng.set('timing', 1)
    m, t = [], []
    for _ in range(10):
        eng.lnlike_batch(tht); torch.cuda.synchronize(); m.append(eng.last_ms('mlp')); t.append(eng.last_ms('tail'))
    print('B=%4d  wall %.1f us   mlp %.1f us   tail %.1f us' % (B, wall * 1e6, np.median(m) * 1e3, np.median(t) * 1e3), flush=True)
