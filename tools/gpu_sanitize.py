"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
for name, fast in [('mini_spec', 1), ('mini_spec', 0), ('mini_joint', 1), ('mini_noinst', 1), ('mini_odd', 1),
                   ('mini_smlp', 1), ('mini_yst', 1), ('c2', 1)]:
    cfg, g = load_case(name)
    eng = engine_from_config(cfg, precision='parity')
    eng.set('fast_tail', fast)
    th = torch.from_numpy(g['theta'][:(3 if name == 'c2' else 6)]).cuda()
    f, m, l = eng.model_batch(th)
    l2 = eng.lnlike_batch(th)
    torch.cuda.synchronize()
    print(name, 'fast' if fast else 'general', 'max|dlnL|', float(np.nanmax(np.abs(l2.cpu().numpy() - g['lnl'][:len(th)]))))
    eng.close()
# round 2: cluster tail (65536- and 32768-sample transforms over four CTAs, distributed shared memory), dynamic point
# scheduling with a capped grid (every CTA claims several points), hidden-layer stack as one cluster launch
for name, nrow in [('c4m', 3), ('mid', 5), ('c2', 7)]:
    cfg, g = load_case(name)
    eng = engine_from_config(cfg, precision='parity')
    eng.set('tail_cluster', 1)
    eng.set('tail_grid_cap', 2)
    th = torch.from_numpy(g['theta'][:nrow]).cuda()
    f, m, l = eng.model_batch(th)
    l2 = eng.lnlike_batch(th)
    torch.cuda.synchronize()
    print(name, 'cluster' if eng.query('tail_cluster') else 'single-CTA', 'capped grid', 'max|dlnL|',
          float(np.nanmax(np.abs(l2.cpu().numpy() - g['lnl'][:len(th)]))))
    eng.close()
