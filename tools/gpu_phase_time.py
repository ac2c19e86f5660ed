"""Tail time at C2 with phases of the fast kernel switched off (dev tool; see TailParams.debug_skip).
Also times batch sizes that are exact multiples of the resident grid (quantisation check)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
eng.set('timing', 1)


def tail_ms(B, mask, n=10):
    eng.set('debug_skip', mask)
    tht = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=1))).cuda()
    for _ in range(3): eng.lnlike_batch(tht)
    t = []
    for _ in range(n):
        eng.lnlike_batch(tht); torch.cuda.synchronize(); t.append(eng.last_ms('tail'))
    return float(np.median(t))


grid = eng.query('tail_grid')
print('tail grid', grid)
for name, mask in [('full', 0), ('no stage 1', 1), ('no stage 2', 2), ('no stages', 3), ('no regrid_back', 8),
                   ('no final', 16), ('only stages', 24), ('nothing', 27)]:
    print('%-18s %.4f ms' % (name, tail_ms(4096, mask)), flush=True)
for B in [grid, 2 * grid, 9 * grid, 4096, 10 * grid]:
    ms = tail_ms(B, 0)
    print('B=%5d  %.4f ms  %.1f ns/point' % (B, ms, ms * 1e6 / B), flush=True)
