import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
th = torch.from_numpy(cfg.draw(4096, seed=2)).cuda()
eng.set('timing', 1)
for fast in (1, 0):
    eng.set('fast_tail', fast)
    for _ in range(3): eng.lnlike_batch(th)
    t = 0
    for _ in range(5):
        eng.lnlike_batch(th); t += eng.last_ms('tail')
    print('fast_tail', fast, 'tail ms', t / 5)
