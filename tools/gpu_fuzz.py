"""Differential fuzz of the CUDA path against the CPU oracle on random small configurations (dev tool).

Random emulator ranges / widths / network types, observed grids that may stick out of the emulator's
coverage, optional continuum polynomial and photometry, parameters drawn from widened boxes (large
vsini, large |vrad|, coarse and too-fine Inst_R).  Prints the worst flux / lnL deviation per config and
exits non-zero on a parity failure.
"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import payne_oracle as O
from thepayne_b200 import synth
from thepayne_b200.engine import engine_from_config

def configs(seed, ncfg, say=lambda *a, **k: None):
    """The fuzz's random stream: yields (index, builder kwargs, config, parameter batch)."""
    rng = np.random.default_rng(seed)
    for it in range(ncfg):
        w0 = float(rng.uniform(4000, 8000))
        span = float(rng.choice([20.0, 45.0, 90.0, 200.0]))
        r_fwhm = float(rng.choice([30000.0, 50000.0, 80000.0]))
        lo = w0 + rng.uniform(-0.05, 0.2) * span          # may start before the emulator -> NaN pixels
        hi = w0 + span * rng.uniform(0.6, 1.05)
        kw = dict(ann_range=(w0, w0 + span), r_fwhm=r_fwhm, obs_range=(lo, hi), n_obs=int(rng.integers(200, 2500)),
                  H=int(rng.choice([16, 40, 64, 100])), vmic=bool(rng.integers(2)), npoly=int(rng.choice([0, 0, 2, 4])),
                  nntype=str(rng.choice(['LinNet', 'LinNet', 'SMLP', 'YST1'])), vrot_max=float(rng.choice([5.0, 60.0, 300.0])),
                  seed_net=int(rng.integers(1000)))
        if rng.integers(3) == 0:
            kw.update(bands=synth.PROCYON_BANDS[:int(rng.integers(2, 7))], photH=16, photscale=bool(rng.integers(2)))
        try:
            cfg = synth.build_config('fuzz%d' % it, model_fn=O.model_fn, **kw)
        except Exception as e:                               # truth outside coverage etc.
            say('cfg %2d skipped at build: %s' % (it, str(e)[:80])); continue
        B = 12
        th = cfg.draw(B, seed=int(rng.integers(1 << 30)))
        ix = {p: i for i, p in enumerate(cfg.fitpars_i)}
        th[1, ix['Vrad']] = float(rng.uniform(-300, 300))
        th[2, ix['Inst_R']] = float(rng.uniform(5000, 20000))
        th[3, ix['Inst_R']] = r_fwhm * 1.3                  # finer than the emulator -> NaN
        th[4, ix['Vrot']] = 0.0
        th[5, ix['Vrad']] = 0.0
        if 'Av' in ix:
            th[6, ix['Av']] = 5.5
        yield it, kw, cfg, th


def run(seed=0, ncfg=24, verbose=True, only=None):
    """Returns the number of configurations that fail parity.  ``only``: evaluate just these configuration
    indices (every configuration is still drawn and built, so the random stream is the same)."""
    bad = 0
    say = print if verbose else (lambda *a, **k: None)
    for it, kw, cfg, th in configs(seed, ncfg, say):
        if only is not None and it not in only:
            continue
        L = O.OracleLikelihood(cfg)
        with np.errstate(all='ignore'):
            ref_l, ref_f, ref_m = L.lnlike_batch(th, return_model=True)
        eng = engine_from_config(cfg, precision='parity')
        flux, mags, lnl = eng.model_batch(torch.from_numpy(np.ascontiguousarray(th)).cuda())
        l2 = eng.lnlike_batch(np.ascontiguousarray(th))
        # the general-grid tail (any increasing emulator grid) on the same points
        lg = None
        if eng.query('nfft1') <= 32768:
            eng.set('fast_tail', 0)
            lg = eng.lnlike_batch(np.ascontiguousarray(th))
            eng.set('fast_tail', 1)
        f = flux.cpu().numpy(); l = lnl.cpu().numpy()
        nan_ok = np.array_equal(np.isnan(f), np.isnan(ref_f)) and np.array_equal(np.isnan(l), np.isnan(ref_l))
        fin = np.isfinite(ref_f)
        df = float(np.max(np.abs(f[fin] - ref_f[fin]) / np.abs(ref_f[fin]))) if fin.any() else 0.0
        ok = np.isfinite(ref_l)
        tol = np.maximum(1e-3, 1e-8 * np.abs(ref_l[ok]))
        dl = np.abs(l[ok] - ref_l[ok]); dl2 = np.abs(l2[ok] - ref_l[ok])
        if lg is not None:
            nan_ok = nan_ok and np.array_equal(np.isnan(lg), np.isnan(ref_l))
            dl2 = np.maximum(dl2, np.abs(lg[ok] - ref_l[ok]))
        worst = float(np.max(np.maximum(dl, dl2) / tol)) if ok.any() else 0.0
        note = ''
        within = worst <= 1.0
        if not within and ok.any():
            # Beyond the bar: is it this implementation or the reference's own fp32 round-off?  Yardstick = the
            # same points with the emulator evaluated in float64 (exact-arithmetic MLP, everything else unchanged).
            Li = O.OracleLikelihood(cfg, ideal_mlp=True)
            idx = np.where(ok)[0][np.maximum(dl, dl2) > tol]
            closer = True
            for i in idx:
                ideal = float(Li.lnlikefn(th[i]))
                e_gpu, e_ref = abs(l[i] - ideal), abs(ref_l[i] - ideal)
                # as close to exact arithmetic as the reference is, or within max(1e-3, 3e-8 |lnL|) of it: at
                # these points (|lnL| 3e4..2.5e5) both implementations sit 1e-4..3e-3 (<= 2.6e-8 relative) from the
                # exact value with either sign.  tools/gpu_fuzz_diag.py splits the deviation: the fp32 tail
                # contributes <= 3e-4, the rest is the float32 rounding of the hidden activations, which every
                # fp32 emulator (the reference's MKL one included) has and no summation order removes
                closer = bool(closer and (e_gpu <= e_ref or e_gpu <= max(1e-3, 3e-8 * abs(ideal))))
                note += '\n      lnL %.3f: gpu-ref %+.2e (rel %.1e) | ref-exact %+.2e, gpu-exact %+.2e' % (
                    ref_l[i], l[i] - ref_l[i], abs(l[i] - ref_l[i]) / abs(ref_l[i]), ref_l[i] - ideal, l[i] - ideal)
            within = closer                                   # at least as close to exact arithmetic as the reference is
            note = ('  [beyond 1e-3 vs the reference; closer to exact arithmetic than the reference, or within max(1e-3, 3e-8 |lnL|) of it]' if closer else '') + note
        good = nan_ok and df < 1e-5 and within and eng.query('status') == 0
        bad += not good
        say('cfg %2d %-6s n_ann %5d n_obs %4d H %3d poly %d phot %d fast %d : flux %.1e  lnL/tol %.2f  nan %s  %s' % (
            it, kw['nntype'], len(cfg.spec.wavelength), kw['n_obs'], kw['H'], kw['npoly'], int(cfg.phot is not None),
            eng.query('fast_tail'), df, worst, nan_ok, ('ok' if good else 'FAIL') + note), flush=True)
        eng.close()
    return bad


if __name__ == '__main__':
    nbad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 24)
    print('failures:', nbad)
    sys.exit(1 if nbad else 0)
