"""Latency of one lnlike_batch call vs batch size (dev tool): device-resident and host entry."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
for B in [1, 16, 64, 128, 256, 512, 1024, 2048, 4096]:
    th = np.ascontiguousarray(cfg.draw(B, seed=B))
    tht = torch.from_numpy(th).cuda()
    for _ in range(5): eng.lnlike_batch(tht)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 30
    for _ in range(n): eng.lnlike_batch(tht)
    torch.cuda.synchronize(); d = (time.perf_counter() - t0) / n
    for _ in range(3): eng.lnlike_batch(th)
    t0 = time.perf_counter()
    for _ in range(n): eng.lnlike_batch(th)
    h = (time.perf_counter() - t0) / n
    print('B=%5d  device %8.1f us (%9.0f evals/s)   host entry %8.1f us (%9.0f evals/s)' % (B, d * 1e6, B / d, h * 1e6, B / h), flush=True)
