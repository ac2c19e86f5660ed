import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import load_case
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='parity')
tht = torch.from_numpy(np.ascontiguousarray(cfg.draw(1, seed=1))).cuda()
for _ in range(4): eng.lnlike_batch(tht)
torch.cuda.synchronize()
