"""Split the lnL error into emulator (MLP) and tail contributions (dev tool).

tail-only error: feed OUR emulator flux (downloaded) through the oracle's fp64 tail and compare
with our lnL; MLP-only error: oracle tail on our flux vs oracle tail on the reference flux.
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from oracle import payne_oracle as O
from thepayne_b200.engine import engine_from_config

name = sys.argv[2] if len(sys.argv) > 2 else 'c2'
prec = sys.argv[1]
cfg, g = load_case(name)
n = 16
th = g['theta'][:n]
eng = engine_from_config(cfg, precision=prec)
_, _, lnl = eng.model_batch(torch.from_numpy(th).cuda())
lnl = lnl.cpu().numpy()
L = O.OracleLikelihood(cfg)
Li = O.OracleLikelihood(cfg, ideal_mlp=True)
x = np.stack([L._col(th, p) for p in ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]']], 1)
y = eng.ann_eval(x).cpu().numpy()
yi = Li.net(x)                     # float64 ideal
yr = np.stack([L.net(r) [0] for r in x])   # reference fp32, batch 1
print('%s %s: emulator flux vs ideal: ours rms %.2e max %.2e mean %.2e | reference rms %.2e max %.2e mean %.2e' % (
    name, prec, np.std(y - yi), np.abs(y - yi).max(), np.mean(y - yi), np.std(yr - yi), np.abs(yr - yi).max(), np.mean(yr - yi)))
print(' row      lnL_ref      ours-ref   tail-only   mlp-only  ref-vs-ideal')
for i in range(n):
    if not np.isfinite(g['lnl'][i]):
        continue
    fo, _ = L.model(th[i], mlp_flux=y[i].astype(np.float32).copy())
    l_tail = -0.5 * np.sum(((fo - cfg.obs_flux) / cfg.obs_eflux) ** 2)
    if L.phot_bool:
        _, mg = L.model(th[i])
        o = np.array([v[0] for v in cfg.obs_phot.values()]); e = np.array([v[1] for v in cfg.obs_phot.values()])
        l_tail += -0.5 * np.sum(((mg - o) / e) ** 2)
    fi, _ = Li.model(th[i])
    l_ideal = -0.5 * np.sum(((fi - cfg.obs_flux) / cfg.obs_eflux) ** 2) + (l_tail * 0 if not L.phot_bool else -0.5 * np.sum(((mg - o) / e) ** 2))
    print('%4d %12.3f  %+.2e  %+.2e  %+.2e  %+.2e' % (i, g['lnl'][i], lnl[i] - g['lnl'][i], lnl[i] - l_tail,
                                                      l_tail - g['lnl'][i], g['lnl'][i] - l_ideal))
