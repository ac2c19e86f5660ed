"""First-contact GPU diagnostic: parity of every golden case + a rough timing (dev tool).

    python tools/gpu_check.py <precision> [case ...]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from conftest import load_case  # noqa: E402
from oracle import goldens, payne_oracle as O  # noqa: E402
from thepayne_b200.engine import engine_from_config  # noqa: E402


def main():
    prec = sys.argv[1]
    names = sys.argv[2:] or list(goldens.CASES)
    print('device', torch.cuda.get_device_name(0), 'precision', prec, flush=True)
    for name in names:
        cfg, g = load_case(name)
        eng = engine_from_config(cfg, precision=prec)
        th = torch.from_numpy(g['theta']).cuda()
        flux, mags, lnl = eng.model_batch(th)
        torch.cuda.synchronize()
        lnl = lnl.cpu().numpy()
        ref = g['lnl']
        nanok = np.array_equal(np.isnan(lnl), np.isnan(ref))
        ok = np.isfinite(ref) & np.isfinite(lnl)
        dl = np.abs(lnl[ok] - ref[ok])
        msg = '%-12s nan-pattern %s  max|dlnL| %.3e (at lnL %.1f)  max rel %.2e' % (
            name, nanok, dl.max() if ok.any() else -1,
            ref[ok][np.argmax(dl)] if ok.any() else 0, (dl / np.abs(ref[ok])).max() if ok.any() else -1)
        if flux is not None and ok.any():
            nf = g['flux'].shape[0]
            f = flux[:nf].cpu().numpy()
            rf = g['flux']
            fin = np.isfinite(rf)
            msg += '  flux nan-eq %s max rel %.2e' % (np.array_equal(np.isnan(f), np.isnan(rf)),
                                                      np.max(np.abs(f[fin] - rf[fin]) / np.abs(rf[fin])))
        if mags is not None:
            m = mags.cpu().numpy()
            msg += '  mags max abs %.2e' % np.max(np.abs(m - g['mags']))
        print(msg, flush=True)
        if name in ('mini_spec', 'c2'):
            # emulator alone vs torch fp32 on the CPU
            L = O.OracleLikelihood(cfg)
            x = np.stack([L._col(g['theta'], p) for p in ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]']], 1)[:8]
            y = eng.ann_eval(x).cpu().numpy()
            yr = L.net(x)
            print('   ann_eval max rel %.2e' % np.max(np.abs(y - yr) / np.abs(yr)), flush=True)
            lh = eng.lnlike_batch(g['theta'])
            ld = eng.lnlike_batch(th).cpu().numpy()          # same lnL-only path (model_batch's lnL goes through the stored spectrum)
            print('   host entry == device entry:', np.array_equal(np.nan_to_num(lh, nan=1.0), np.nan_to_num(ld, nan=1.0)),
                  ' max|lnL-only - via spectrum| %.2e' % np.nanmax(np.abs(ld - lnl)))
        if name == 'c2':
            B = 4096
            thb = torch.from_numpy(cfg.draw(B)).cuda()
            eng.set('timing', 1)
            for _ in range(3):
                out = eng.lnlike_batch(thb)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                out = eng.lnlike_batch(thb)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            print('   C2 B=%d: %.3f ms/batch = %.3e evals/s ; mlp %.3f ms tail %.3f ms ; finite %d' % (
                B, dt * 1e3, B / dt, eng.last_ms('mlp'), eng.last_ms('tail'), int(torch.isfinite(out).sum())),
                flush=True)
            print('   status flag', eng.query('status'), 'launches', eng.query('launches'))
        eng.close()


if __name__ == '__main__':
    main()
