"""Tail time at C2 against the vsini range of the batch (does the rotation-table window cover it?) -- dev tool."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
eng.set('timing', 1)
print('rot_window_floats', eng.query('rot_window_floats'))
iv = cfg.fitpars_i.index('Vrot')
for vmax in [5.0, 4.5, 4.0, 3.0, 1.0]:
    th = cfg.draw(4096, seed=1)
    th[:, iv] *= vmax / 5.0
    tht = torch.from_numpy(np.ascontiguousarray(th)).cuda()
    for _ in range(3): eng.lnlike_batch(tht)
    t = []
    for _ in range(10):
        eng.lnlike_batch(tht); torch.cuda.synchronize(); t.append(eng.last_ms('tail'))
    print('vrot <= %.1f : tail %.4f ms' % (vmax, float(np.median(t))), flush=True)
