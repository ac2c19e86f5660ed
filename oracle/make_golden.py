"""Mint tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden [case ...]

For every case in ``oracle/goldens.py`` the reference's own
``likelihood.lnlikefn -> genspec -> getspec -> smoothspec -> chi2`` is executed through the
stub-import harness (``oracle/refharness.py``) and the inputs/outputs are stored:
theta, the mock observation, reference model flux (first rows), magnitudes and lnL.
"""
import os
import sys

import numpy as np

from oracle import goldens, refharness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main(names):
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        # the demo file lives in the reference checkout (not among the staged hot-path files)
        data = goldens.demo_observation(os.environ.get('PAYNE_REFERENCE_CHECKOUT', '/root/reference')) if name == 'c1' else None
        cfg = goldens.build(name, refharness.ref_model_fn, data=data)
        th = goldens.thetas(name, cfg)
        nflux = goldens.CASES[name][2]
        lnl = refharness.ref_lnlike(cfg, th)
        fl, mg = refharness.ref_model_fn(cfg, th[:nflux])
        # ref_model_fn swaps the observation out; mags need all rows
        _, mg = refharness.ref_model_fn(cfg, th) if cfg.phot is not None else (None, None)
        d = dict(theta=th, lnl=lnl, obs_wave=cfg.obs_wave, obs_flux=cfg.obs_flux,
                 obs_eflux=cfg.obs_eflux, fitpars=np.array(cfg.fitpars_i),
                 digest=np.array(cfg.spec.digest()), flux=fl)
        if mg is not None:
            d['mags'] = mg
            d['obs_phot'] = np.array([cfg.obs_phot[b] for b in cfg.phot.bands])
        np.savez_compressed(os.path.join(OUT, name + '.npz'), **d)
        print(name, 'ndim', cfg.ndim, 'points', len(th), 'lnl[:4]', lnl[:4],
              'nan', int(np.isnan(lnl).sum()))


def prior_golden():
    """Unit-cube samples -> reference prior.priortrans / lnpriorfn (Payne/fitting/prior.py)."""
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    from test_batching import FREE, NAMES, PRIORS
    U = np.random.default_rng(0).random((64, len(FREE)))
    U[0, :] = 1.0 - 1e-16
    theta, lnp = refharness.ref_prior(PRIORS, NAMES, FREE, [True, True, True, True, False], U)
    np.savez_compressed(os.path.join(OUT, 'prior.npz'), U=U, theta=theta, lnp=lnp)
    print('prior', theta.shape, 'finite', int(np.isfinite(theta).all()), 'lnp[:3]', lnp[:3])


def getspec_golden():
    """``f3_getspec``: the reference's getspec with a continuum emulator / an LSF vector."""
    spec, cont, calls = goldens.getspec_case()
    d = dict(digest=np.array(spec.digest()), cdigest=np.array(cont.digest()))
    for i, kw in enumerate(calls):
        kw = dict(kw)
        use_cont = kw.pop('use_cont')
        (w, f), = refharness.ref_getspec(spec, cont if use_cont else None, [kw])
        d['wave_%d' % i], d['flux_%d' % i] = w, f
        print('f3_getspec call', i, 'n', len(f), 'nan', int(np.isnan(f).sum()), 'mean', np.nanmean(f))
    np.savez_compressed(os.path.join(OUT, 'f3_getspec.npz'), **d)


if __name__ == '__main__':
    names = sys.argv[1:] or (list(goldens.CASES) + ['prior', 'f3_getspec'])
    if 'f3_getspec' in names:
        names.remove('f3_getspec')
        getspec_golden()
    if 'prior' in names:
        names.remove('prior')
        prior_golden()
    main(names)
