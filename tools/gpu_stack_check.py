"""Dev tool: emulator output of several cases / batch sizes saved to gpurun_out/stack_<tag>.npz; run once with
PAYNE_GEMM_STACK=0 and once without, then compare bit for bit with `gpu_stack_check.py compare`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = os.path.join(ROOT, 'gpurun_out')
if len(sys.argv) > 1 and sys.argv[1] == 'compare':
    a, b = np.load(os.path.join(OUT, 'stack_0.npz')), np.load(os.path.join(OUT, 'stack_1.npz'))
    for k in a.files:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        print(k, a[k].shape, 'bit-identical' if same else 'DIFFERENT max|d| %.3e' % np.nanmax(np.abs(a[k] - b[k])))
    sys.exit(0)
import torch
from conftest import load_case
from thepayne_b200.engine import engine_from_config
tag = '0' if os.environ.get('PAYNE_GEMM_STACK') == '0' else '1'
res = {}
for name in ('mini_spec', 'c2', 'c4m', 'mini_odd'):
    cfg, g = load_case(name)
    eng = engine_from_config(cfg, precision='parity')
    for B in (1, 5, 130, 1000):
        x = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=B)[:, :eng.D_in])).cuda()
        res['%s_%d' % (name, B)] = eng.ann_eval(x).cpu().numpy()
    th = torch.from_numpy(np.ascontiguousarray(cfg.draw(300, seed=2))).cuda()
    res[name + '_lnl'] = eng.lnlike_batch(th).cpu().numpy()
    eng.close()
np.savez(os.path.join(OUT, 'stack_%s.npz' % tag), **res)
print('saved', tag, len(res))
