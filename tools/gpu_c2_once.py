"""C2 batch through the hot path a few times (target for ncu captures) -- dev tool.  usage: gpu_c2_once.py [B] [reps] [case]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg, g = load_case(sys.argv[3] if len(sys.argv) > 3 else 'c2')
eng = engine_from_config(cfg, precision='parity')
th = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=1))).cuda()
for _ in range(reps):
    out = eng.lnlike_batch(th)
torch.cuda.synchronize()
print('finite', int(torch.isfinite(out).sum()))
