"""GPU: the two getspec branches the likelihood never reaches -- continuum emulator
(predictspec.py:96-102, 208-226) and LSF-vector broadening (predictspec.py:265-286 -> smoothing.py:482-586)
-- through the reference-named mirror, against a fixture minted by the unmodified reference
(tests/golden/f3_getspec.npz, oracle/make_golden.py) and against the CPU oracle on other inputs."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import goldens, payne_oracle as O

pytestmark = pytest.mark.gpu
FLUX_BAR = 1e-5          # north-star per-pixel bar; measured values are asserted against a tighter guard too


def _predictor(spec, cont, prec='parity'):
    from thepayne_b200.predict.predictspec import PayneSpecPredict
    return PayneSpecPredict(nnpath=spec, Cnnpath=cont, precision=prec)


@pytest.mark.parametrize('prec', ['parity', 'simt'])
def test_getspec_continuum_and_lsf_vs_reference(prec):
    g = np.load(os.path.join(GOLDEN, 'f3_getspec.npz'))
    spec, cont, calls = goldens.getspec_case()
    assert spec.digest() == str(g['digest']) and cont.digest() == str(g['cdigest'])
    pp = {True: _predictor(spec, cont, prec), False: _predictor(spec, None, prec)}
    for i, kw in enumerate(calls):
        kw = dict(kw)
        P = pp[kw.pop('use_cont')]
        w, f = P.getspec(**kw)
        ref = g['flux_%d' % i]
        np.testing.assert_array_equal(w, g['wave_%d' % i])
        assert f.shape == ref.shape
        bad = np.isnan(f) != np.isnan(ref)
        # outwave=None with a scalar inst_R: whether the two end pixels of the native grid fall inside the
        # resampled grid is decided by the last bit of exp(log(w)) in the reference (smoothing.py:289)
        assert bad.sum() == 0 or (kw['outwave'] is None and set(np.flatnonzero(bad)) <= {0, len(f) - 1}), (i, bad.sum())
        ok = np.isfinite(f) & np.isfinite(ref)
        err = np.max(np.abs(f[ok] - ref[ok]) / np.abs(ref[ok]))
        assert err <= FLUX_BAR and err <= 5e-7, (i, err)


def test_lsf_batch_vs_oracle_and_switching():
    """Batched LSF spectra against the oracle at other labels / velocities; detaching the vector gives the
    scalar-R stage back; a vector of the wrong length is refused."""
    spec, cont, calls = goldens.getspec_case()
    outwave, lsf = calls[3]['outwave'], calls[3]['inst_R']
    P = _predictor(spec, None)
    rng = np.random.default_rng(3)
    B = 6
    labels = np.stack([rng.uniform(4200, 7500, B), rng.uniform(4.0, 5.0, B), rng.uniform(-0.1, 0.1, B),
                       rng.uniform(-0.1, 0.1, B)], axis=1)
    rot = np.array([0.0, 1.0, 4.0, 9.0, 20.0, 2.5])
    rad = np.array([0.0, -80.0, 35.0, 0.0, 150.0, -3.0])
    f = P.getspec_batch(labels, rot, rad, np.full(B, np.nan), outwave, lsf=lsf).cpu().numpy()
    net = O.make_net(spec)
    for b in range(B):
        _, fr = O.getspec(net, spec, *labels[b], np.nan, rot[b], rad[b], lsf, outwave)
        assert np.array_equal(np.isnan(f[b]), np.isnan(fr))
        assert np.max(np.abs(f[b] - fr) / np.abs(fr)) <= 5e-7
    eng = P.anns.engine_for(outwave, inst_sigma=True, lsf=lsf)
    assert eng.query('lsf') == 1 and eng.query('status') == 0
    eng.set_lsf(None)
    assert eng.query('lsf') == 0
    th = np.array([[5770.0, 4.44, 0.0, 0.0, 0.5, 3.0, np.nan, 32000.0 * 2.355]])
    fs, _, _ = eng.model_batch(th, want_mags=False)
    _, fr = O.getspec(net, spec, 5770.0, 4.44, 0.0, 0.0, np.nan, 3.0, 0.5, 32000.0 * 2.355, outwave)
    assert np.max(np.abs(fs[0].cpu().numpy() - fr) / np.abs(fr)) <= 5e-7
    from thepayne_b200._lib import PayneError
    with pytest.raises((PayneError, ValueError)):
        eng.set_lsf(lsf[:-1])
    with pytest.raises(AssertionError):
        P.getspec(Teff=5770.0, inst_R=lsf[:-3], outwave=outwave)


def test_continuum_lnl_path_and_vmic():
    """A continuum emulator attached to a likelihood engine (5-label nets, polynomial, photometry off):
    lnL and spectra against the oracle's getspec + polycalc + chi2."""
    import torch
    from thepayne_b200 import synth
    from thepayne_b200.engine import engine_from_config
    cfg = synth.config_mini(O.model_fn, vmic=True, npoly=2)
    cwave = np.linspace(5138.0, 5195.0, 257)
    cont = synth.make_specnet(5, 24, cwave, cfg.spec.resolution, seed=9)
    th = cfg.theta_true + 0.15 * (cfg.draw(6, seed=8) - cfg.theta_true)   # |lnL| of a few 1e3: the flat 1e-3 bar applies
    th[0] = cfg.theta_true
    eng = engine_from_config(cfg)
    eng.attach_continuum(cont)
    assert eng.query('continuum') == 1
    flux, _, lnl = eng.model_batch(torch.from_numpy(th).cuda(), want_mags=False)
    flux, lnl = flux.cpu().numpy(), lnl.cpu().numpy()
    ix = {p: i for i, p in enumerate(cfg.fitpars_i)}

    def oracle(ideal):
        net, cnet = O.make_net(cfg.spec, ideal=ideal), O.make_net(cont, ideal=ideal)
        fl, ll = [], []
        for t in th:
            _, fr = O.getspec(net, cfg.spec, t[ix['Teff']], t[ix['log(g)']], t[ix['[Fe/H]']], t[ix['[a/Fe]']], t[ix['Vmic']],
                              t[ix['Vrot']], t[ix['Vrad']], 2.355 * t[ix['Inst_R']], cfg.obs_wave, cont=(cnet, cont))
            fr = fr * O.polycalc([t[ix['pc_0']], t[ix['pc_1']]], cfg.obs_wave)
            fl.append(fr)
            ll.append(-0.5 * np.sum(((fr - cfg.obs_flux) ** 2.0) / (cfg.obs_eflux ** 2.0)))
        return np.array(fl), np.array(ll)
    fr, lr = oracle(False)          # the reference's arithmetic: both emulators in torch float32
    fi, li = oracle(True)           # both emulators in float64
    assert np.max(np.abs(flux - fr) / np.abs(fr)) <= 5e-7
    # The continuum net's float32 output (1 ulp = 6e-8 of a value near 1) is interpolated from a grid ~6x coarser
    # than the observed one, so its rounding is coherent over several observed pixels and moves lnL by a few
    # 1e-3 even at |lnL| ~ 1e4 -- in the reference just as here.  The bar is therefore the reference's own
    # distance from exact arithmetic on these rows (never below the flat 1e-3).
    floor = max(1e-3, 2.0 * float(np.max(np.abs(lr - li))))
    assert np.all(np.abs(lnl - lr) <= floor) and np.all(np.abs(lnl - li) <= floor), (lnl - lr, lr - li)
    eng.close()
