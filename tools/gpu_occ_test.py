import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
eng.set('timing', 1)
def tail_ms(B, mask, n=15):
    eng.set('debug_skip', mask)
    tht = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=1))).cuda()
    for _ in range(3): eng.lnlike_batch(tht)
    t = []
    for _ in range(n):
        eng.lnlike_batch(tht); torch.cuda.synchronize(); t.append(eng.last_ms('tail'))
    return float(np.median(t))
for mask, name in [(0, 'full'), (28, 'only ffts'), (3, 'no ffts')]:
    for B in [1, 74, 148, 296, 444]:
        print(name, 'B=%d  %.2f us' % (B, tail_ms(B, mask) * 1e3), flush=True)
