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
