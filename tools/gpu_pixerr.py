import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from oracle import payne_oracle as O
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
eng = engine_from_config(cfg, precision='simt')
L = O.OracleLikelihood(cfg)
base = g['theta'][13].copy()
ix = {p: i for i, p in enumerate(cfg.fitpars_i)}
rows = []
for v in [0.0, 0.05, 0.3, 0.735, 1.5, 3.0, 6.0, 12.0]:
    r = base.copy(); r[ix['Vrot']] = v; rows.append(r)
r = g['theta'][7].copy(); rows.append(r)
th = np.array(rows)
flux, _, lnl = eng.model_batch(torch.from_numpy(th).cuda())
flux = flux.cpu().numpy()
x = th[:, :4]
y = eng.ann_eval(x).cpu().numpy()
for i in range(len(th)):
    fo, _ = L.model(th[i], mlp_flux=y[i].astype(np.float32).copy())
    d = flux[i] - fo
    j = np.argmax(np.abs(d))
    lt = -0.5 * np.sum(((fo - cfg.obs_flux) / cfg.obs_eflux) ** 2)
    print('vrot %6.3f vrad %6.2f: dlnL %+.2e  flux err rms %.2e max %.2e at pix %d  mean %+.2e  corr(d, resid) %.3f' % (
        th[i, ix['Vrot']], th[i, ix['Vrad']], lnl[i].item() - lt, d.std(), np.abs(d).max(), j, d.mean(),
        np.corrcoef(d, fo - cfg.obs_flux)[0, 1]))
    if i in (3, 8):
        print('    err[::700]', np.array2string(d[::700], precision=2))
        k = np.argsort(-np.abs(d))[:8]; print('    worst pixels', sorted(k.tolist()))
