"""Where does the lnL error of one fuzz configuration come from (dev tool)?  usage: gpu_fuzz_diag.py seed cfg
Per finite row: ours - reference, tail-only (oracle fp64 tail on OUR emulator flux vs ours), emulator-only,
reference vs exact arithmetic; and the per-pixel error of our model spectrum regressed on (depth, d/dx, d2/dx2, 1)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools'))
import gpu_fuzz
from oracle import payne_oracle as O
from thepayne_b200.engine import engine_from_config
seed, want = int(sys.argv[1]), int(sys.argv[2])
for it, kw, cfg, th in gpu_fuzz.configs(seed, want + 1):
    if it != want:
        continue
    L, Li = O.OracleLikelihood(cfg), O.OracleLikelihood(cfg, ideal_mlp=True)
    with np.errstate(all='ignore'):
        ref_l, ref_f, _ = L.lnlike_batch(th, return_model=True)
    labs = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]'] + (['Vmic'] if kw['vmic'] else [])
    x = np.stack([L._col(th, p) for p in labs], 1)
    for prec in ['parity', 'simt']:
        for fast in [1, 0]:
            eng = engine_from_config(cfg, precision=prec)
            eng.set('fast_tail', fast)
            flux, _, lnl = eng.model_batch(torch.from_numpy(np.ascontiguousarray(th)).cuda())
            flux, lnl = flux.cpu().numpy(), lnl.cpu().numpy()
            y = eng.ann_eval(x).cpu().numpy()
            print('== %s fast_tail %d  n_ann %d n_obs %d npoly %d' % (prec, fast, cfg.spec.D_out, len(cfg.obs_wave), kw['npoly']))
            print(' row      lnL_ref   ours-ref  tail-only   mlp-only ref-vs-exact ours-vs-exact |  e rms   amp->dlnL  shift->dlnL width->dlnL const->dlnL rest')
            for i in range(len(th)):
                if not np.isfinite(ref_l[i]):
                    continue
                fo, _ = L.model(th[i], mlp_flux=y[i].astype(np.float32).copy())     # fp64 tail on OUR emulator flux
                s2 = 1.0 / cfg.obs_eflux ** 2
                l_tail = -0.5 * np.sum((fo - cfg.obs_flux) ** 2 * s2)
                fi, _ = Li.model(th[i])
                l_ideal = -0.5 * np.sum((fi - cfg.obs_flux) ** 2 * s2)
                e = flux[i] - fo
                resid = fo - cfg.obs_flux
                b = np.stack([fo - 1.0, np.gradient(fo), np.gradient(np.gradient(fo)), np.ones_like(fo)], 1)
                c, *_ = np.linalg.lstsq(b, e, rcond=None)
                parts = [-np.sum(resid * (c[k] * b[:, k]) * s2) for k in range(4)]
                rest = -np.sum(resid * (e - b @ c) * s2)
                print('%4d %12.3f  %+.2e  %+.2e  %+.2e  %+.2e  %+.2e | %.1e  %+.2e  %+.2e  %+.2e  %+.2e  %+.2e' % (
                    i, ref_l[i], lnl[i] - ref_l[i], lnl[i] - l_tail, l_tail - ref_l[i], ref_l[i] - l_ideal, lnl[i] - l_ideal,
                    e.std(), parts[0], parts[1], parts[2], parts[3], rest))
            eng.close()
