import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from oracle import payne_oracle as O
from thepayne_b200.engine import engine_from_config
cfg, g = load_case('c2')
L = O.OracleLikelihood(cfg)
ix = {p: i for i, p in enumerate(cfg.fitpars_i)}
rows = []
for vr, R in [(0.0, np.nan), (1.3, np.nan), (0.0, 32653.0), (1.3, 32653.0), (6.0, np.nan), (6.0, 32653.0)]:
    r = g['theta'][7].copy(); r[ix['Vrot']] = vr; r[ix['Inst_R']] = R; rows.append(r)
th = np.array(rows)
for fast in (1, 0):
    eng = engine_from_config(cfg, precision='simt')
    eng.set('fast_tail', fast)
    flux, _, lnl = eng.model_batch(torch.from_numpy(th).cuda())
    flux = flux.cpu().numpy()
    y = eng.ann_eval(th[:, :4]).cpu().numpy()
    print('fast_tail', fast)
    for i in range(len(th)):
        fo, _ = L.model(th[i], mlp_flux=y[i].astype(np.float32).copy())
        e = flux[i] - fo
        resid = fo - cfg.obs_flux
        d = fo - 1.0
        b = np.stack([d, np.gradient(fo), np.gradient(np.gradient(fo)), np.ones_like(fo)], 1)
        c, *_ = np.linalg.lstsq(b, e, rcond=None)
        s2 = 1.0 / cfg.obs_eflux ** 2
        dl_tot = -np.sum(resid * e * s2)
        parts = [-np.sum(resid * (c[k] * b[:, k]) * s2) for k in range(4)]
        print(' R %7.0f vrot %8.1e vrad %6.1f dlnL(lin) %+.2e | amp c=%+.2e -> %+.2e | shift c=%+.2e px -> %+.2e | width c=%+.2e -> %+.2e | const c=%+.2e -> %+.2e | rms e %.2e resid-fit %.2e' % (
            th[i, ix['Inst_R']], th[i, ix['Vrot']], th[i, ix['Vrad']], dl_tot, c[0], parts[0], c[1], parts[1], c[2], parts[2], c[3], parts[3],
            e.std(), (e - b @ c).std()))
    eng.close()
