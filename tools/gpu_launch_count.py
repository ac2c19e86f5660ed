"""Dev tool: kernel launches per likelihood call (9 = one launch per hidden layer, 6 = hidden layers in one launch) and
B = 1 latency."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case(sys.argv[1] if len(sys.argv) > 1 else 'c2')
eng = engine_from_config(cfg, precision='parity')
for B in (1, 128, 4096):
    th = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=1))).cuda()
    for _ in range(5): eng.lnlike_batch(th)
    torch.cuda.synchronize()
    l0 = eng.query('launches')
    eng.lnlike_batch(th); torch.cuda.synchronize()
    l1 = eng.query('launches')
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    ev0.record()
    for _ in range(n): eng.lnlike_batch(th)
    ev1.record(); torch.cuda.synchronize()
    thh = th.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(n): eng.lnlike_batch(thh)
    dt = (time.perf_counter() - t0) / n
    print('B=%d: %d launches per call, %.1f us per call back to back on the device, %.1f us through the host entry' % (
        B, l1 - l0, 1e3 * ev0.elapsed_time(ev1) / n, 1e6 * dt), flush=True)
