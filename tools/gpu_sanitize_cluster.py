"""compute-sanitizer target: the cluster tail alone on two points (dev tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case(sys.argv[1] if len(sys.argv) > 1 else 'mid')
eng = engine_from_config(cfg, precision='parity')
eng.set('tail_cluster', 1)
eng.set('tail_grid_cap', 1)
th = torch.from_numpy(g['theta'][:2]).cuda()
l2 = eng.lnlike_batch(th)
torch.cuda.synchronize()
print('max|dlnL|', float(np.nanmax(np.abs(l2.cpu().numpy() - g['lnl'][:2]))))
