"""Dev tool: device time of the MLP and tail stages of one case (default C2, B = 4096) over many steps; run it under
different PAYNE_LIB_PATH values for an A/B of two builds on the same box.  usage: gpu_tail_time.py [case] [B] [steps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
cfg, g = load_case(name)
eng = engine_from_config(cfg, precision='parity')
th = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=1))).cuda()
for _ in range(10): eng.lnlike_batch(th)
torch.cuda.synchronize()
eng.set('timing', 1)
tail, mlp = [], []
for _ in range(steps):
    eng.lnlike_batch(th); torch.cuda.synchronize()
    tail.append(eng.last_ms('tail')); mlp.append(eng.last_ms('mlp'))
print('%s %s B=%d: tail median %.4f ms (min %.4f), mlp median %.4f ms, window %d floats' % (
    os.environ.get('PAYNE_LIB_PATH', 'default'), name, B, np.median(tail), np.min(tail), np.median(mlp), eng.query('rot_window_floats')))
