"""Sampler-side batching: vectorised prior transforms equal the scalar ones,
the pool adapter turns a proposal queue into one batch, and -- on a GPU -- a stand-in nested
sampler driven through the pool reproduces the scalar ``lnprobfn`` values."""
import numpy as np
import pytest

PRIORS = {
    'Teff': {'pv_uniform': [4000.0, 8000.0]}, 'log(g)': {'pv_gaussian': [4.4, 0.2]},
    '[Fe/H]': {'pv_tgaussian': [-0.5, 0.5, 0.0, 0.1]},
    '[a/Fe]': {'pv_exp': [-0.1, 0.05]}, 'Vrad': {'pv_texp': [-1.0, 1.0, 0.5]},
    'Vrot': {'uniform': [0.5, 250.0]},
    'Inst_R': {'pv_tgaussian': [30000.0, 37000.0, 32000.0, 1000.0], 'gaussian': [32000.0, 1500.0]},
    'log(A)': {'pv_uniform': [-3.0, 7.0]}, 'Av': {'pv_uniform': [0.0, 1.0], 'gaussian': [0.5, 0.2]},
    'blaze_coeff': [[0.0, 1.0], [0.0, 0.02], [0.0, 0.02]],
}
NAMES = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'log(A)', 'Av',
         'Rv', 'CarbonScale', 'pc_0', 'pc_1', 'pc_2']
FREE = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Inst_R', 'log(A)', 'Av', 'pc_0', 'pc_1', 'pc_2']


def make_prior():
    from thepayne_b200.fitting.prior import prior
    flags = {n: n in FREE for n in NAMES}
    return prior({'fixedpars': {}}, PRIORS, [NAMES, flags], [True, True, True, True, False])


def test_prior_batch_equals_scalar():
    P = make_prior()
    assert P.fitpars_i == FREE
    U = np.random.default_rng(0).random((64, P.ndim))
    U[0, :] = 1.0 - 1e-16          # upper edge: truncated priors clamp instead of returning inf
    TB = P.priortrans_batch(U)
    TS = np.array([P.priortrans(list(u)) for u in U])
    np.testing.assert_array_equal(TB, TS)
    assert np.all(np.isfinite(TB))
    assert TB[:, 0].min() >= 4000 and TB[:, 0].max() <= 8000 and TB[:, 6].max() <= 37000
    lb = P.lnprior_batch(TB)
    ls = np.array([P.lnpriorfn({k: v for k, v in zip(FREE, t)}) for t in TB])
    np.testing.assert_allclose(lb, ls, rtol=1e-15)
    TB[3, FREE.index('Vrot')] = 0.1          # outside the additive uniform prior
    assert P.lnprior_batch(TB)[3] == -np.inf and np.isfinite(P.lnprior_batch(TB)[4])


def test_prior_matches_reference_golden():
    """tests/golden/prior.npz: the UNMODIFIED reference prior.priortrans / lnpriorfn on the same
    unit-cube sample (oracle/make_golden.py prior)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'prior.npz'))
    P = make_prior()
    np.testing.assert_allclose(P.priortrans_batch(g['U']), g['theta'], rtol=1e-14, atol=0)
    np.testing.assert_allclose(P.lnprior_batch(g['theta']), g['lnp'], rtol=1e-13, atol=1e-13)


class _FakeLike:
    """CPU stand-in with the likelihood's batched interface (no GPU needed)."""
    def __init__(self):
        self.calls, self.parsdict = [], {}

    def lnlikefn(self, pars):
        self.parsdict = dict(zip(FREE, pars))
        return float(-0.5 * np.sum(np.asarray(pars) ** 2) * 1e-6)

    def lnlike_batch(self, theta):
        self.calls.append(len(theta))
        return -0.5 * np.sum(theta ** 2, axis=1) * 1e-6


def test_pool_batches_queue():
    from thepayne_b200.fitting.batching import BatchedLnProb, BatchedPool
    P, L = make_prior(), _FakeLike()
    f = BatchedLnProb(L, P)
    pool = BatchedPool(queue_size=32)
    U = np.random.default_rng(1).random((32, P.ndim))
    pts = list(P.priortrans_batch(U))

    class Wrapper:                      # dynesty wraps the callable before mapping it
        def __init__(self, func):
            self.func = func

        def __call__(self, x):
            return self.func(x)
    out = pool.map(Wrapper(f), pts)
    assert L.calls == [32] and pool.batches == 1 and pool.size == 32
    ref = [f(list(p)) for p in pts]
    np.testing.assert_allclose(out, ref, rtol=1e-14)
    assert pool.map(lambda x: 2 * x, [1, 2, 3]) == [2, 4, 6]


@pytest.mark.gpu
def test_standin_nested_sampler_through_the_pool():
    """A minimal nested-sampling loop (replace the worst live point by a better prior draw) whose
    likelihood calls all go through BatchedPool -> likelihood.lnlike_batch on the GPU."""
    from conftest import load_case
    from thepayne_b200.fitting.batching import BatchedLnProb, BatchedPool
    from thepayne_b200.fitting.likelihood import likelihood
    from thepayne_b200.fitting.prior import prior
    cfg, g = load_case('mini_spec')
    names = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R', 'log(R)', 'Dist', 'log(A)',
             'Av', 'Rv', 'CarbonScale']
    flags = {n: n in cfg.fitpars_i for n in names}
    fitargs = {'obs_wave_fit': cfg.obs_wave, 'obs_flux_fit': cfg.obs_flux, 'obs_eflux_fit': cfg.obs_eflux,
               'specANNpath': cfg.spec, 'NNtype': 'LinNet', 'fixedpars': {}}
    pri = {p: {'pv_uniform': list(cfg.box[p])} for p in cfg.fitpars_i}
    L = likelihood(fitargs, [names, flags], cfg.runbools)
    P = prior(fitargs, pri, [names, flags], cfg.runbools)
    f = BatchedLnProb(L, P)
    pool = BatchedPool(queue_size=64)
    rng = np.random.default_rng(3)
    live_u = rng.random((64, P.ndim))
    live = P.priortrans_batch(live_u)
    lnl = np.array(pool.map(f, list(live)))
    np.testing.assert_allclose(lnl[:4], [f(list(t)) for t in live[:4]], rtol=0, atol=1e-6)
    first = lnl.min()
    for _ in range(6):
        cand = P.priortrans_batch(rng.random((pool.size, P.ndim)))
        cl = np.array(pool.map(f, list(cand)))
        for c, l in zip(cand, cl):
            w = np.argmin(lnl)
            if l > lnl[w]:
                live[w], lnl[w] = c, l
    assert pool.batches == 7 and pool.points == 64 * 7
    assert lnl.min() > first and np.isfinite(lnl).all()


@pytest.mark.gpu
def test_fitpayne_batched_fit_recovers_the_mock(tmp_path):
    """``FitPayne.run(inputdict=...)`` (the reference's user entry, fitstar.py:19-217) on a joint
    spectrum + photometry mock with a continuum polynomial, sampled by the lock-step batched nested
    sampler: >= 95 % of the likelihood evaluations go through calls of at least ``queue_size`` points,
    the truth lies within the posterior, and the output file has the reference's layout."""
    from conftest import load_case
    from thepayne_b200.fitting.fitstar import FitPayne
    cfg, g = load_case('mini_joint')
    Q = 128
    inputdict = {
        'spec': {'obs_wave': cfg.obs_wave, 'obs_flux': cfg.obs_flux, 'obs_eflux': cfg.obs_eflux,
                 'normspec': False, 'convertair': False, 'modpoly': True},
        'specANNpath': cfg.spec, 'NNtype': 'LinNet',
        'phot': dict(cfg.obs_phot), 'photANNpath': cfg.phot, 'photscale': True,
        'sampler': {'samplertype': 'Static', 'samplemethod': 'rwalk', 'npoints': Q, 'walks': 25,
                    'delta_logz_final': 0.5, 'flushnum': 100, 'seed': 11},
        'priordict': {p: {'pv_uniform': list(cfg.box[p])} for p in cfg.fitpars_i if not p.startswith('pc_')},
        'output': str(tmp_path / 'demoout.dat'),
    }
    inputdict['priordict']['Inst_R'] = {'pv_tgaussian': [30000.0, 37000.0, 32000.0, 1000.0]}   # demo/runPayne.py:137-139
    inputdict['priordict']['blaze_coeff'] = [[0.0, 1.0], [0.0, 0.02], [0.0, 0.02]]
    FS = FitPayne()
    sampler = FS.run(inputdict=inputdict, verbose=False)
    assert FS.likeobj.fitpars_i == cfg.fitpars_i
    r = sampler.results
    b = r['batch_sizes']
    assert b.sum() == r['ncall']
    assert b[b >= Q].sum() / b.sum() >= 0.95, b[b >= Q].sum() / b.sum()
    assert np.isfinite(r['logz']) and r['niter'] > 10 * Q
    mean, sd = sampler.posterior_mean_std()
    z = (mean - cfg.theta_true) / sd
    assert np.all(sd > 0) and np.all(np.abs(z) < 5.0), dict(zip(cfg.fitpars_i, np.round(z, 2)))
    # the best point beats the truth's likelihood only by the usual ~ndim/2
    lnl_true = float(FS.likeobj.lnlike_batch(cfg.theta_true[None, :])[0])
    assert r['logl'].max() > lnl_true - 1.0 and r['logl'].max() < lnl_true + 40.0
    rows = open(inputdict['output']).read().strip().split('\n')
    assert rows[0].split()[:3] == ['Iter', 'Teff', 'log(g)'] and rows[0].split()[-1] == 'delta(log(z))'
    assert len(rows) == 1 + r['niter'] + Q
    assert len(rows[1].split()) == 1 + len(cfg.fitpars_i) + 7
