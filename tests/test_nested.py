"""CPU: the lock-step batched nested sampler (thepayne_b200/fitting/nested.py) on problems with
known evidence, its batching bookkeeping, and what the older pool shim does when a dynesty-like
sampler maps proposal evolutions (not likelihood calls) through it."""
import math

import numpy as np

from thepayne_b200.fitting.batching import BatchedLnProb, BatchedPool
from thepayne_b200.fitting.nested import BatchedNestedSampler


def _gauss_problem(ndim, sigma, centre=0.5):
    def lnprob(theta):
        r = (np.asarray(theta) - centre) / sigma
        return -0.5 * np.sum(r * r, axis=1) - ndim * math.log(sigma * math.sqrt(2 * math.pi))
    return lnprob, (lambda u: np.asarray(u, dtype=float))       # unit-cube prior; logZ = 0 (mass inside the cube)


def test_gaussian_evidence_and_posterior():
    ndim, sigma = 4, 0.05
    lnprob, pt = _gauss_problem(ndim, sigma)
    s = BatchedNestedSampler(lnprob, pt, ndim, nlive=200, walks=25, seed=1)
    r = s.run_nested(dlogz=0.05)
    assert abs(r['logz']) < 4.0 * max(r['logzerr'], 0.1), (r['logz'], r['logzerr'])
    m, sd = s.posterior_mean_std()
    assert np.all(np.abs(m - 0.5) < 0.2 * sigma * 3), m
    assert np.all(np.abs(sd / sigma - 1.0) < 0.2), sd
    # information of a Gaussian posterior in a unit box: H = -n (ln(sigma sqrt(2 pi e)))
    h_true = -ndim * math.log(sigma * math.sqrt(2 * math.pi * math.e))
    assert abs(r['h'] - h_true) < 0.15 * h_true
    assert abs(r['weights'].sum() - 1.0) < 1e-12
    assert np.all(np.diff(r['logl']) >= 0)          # dead points leave in order of likelihood


def test_every_call_is_a_full_batch():
    ndim = 3
    lnprob, pt = _gauss_problem(ndim, 0.1)
    calls = []

    def counted(theta):
        calls.append(len(theta))
        return lnprob(theta)
    Q = 64
    s = BatchedNestedSampler(counted, pt, ndim, nlive=100, walks=20, queue_size=Q, seed=3)
    s.run_nested(dlogz=0.1)
    calls = np.array(calls)
    assert calls.sum() == s.ncall == s.results['batch_sizes'].sum()
    frac = calls[calls >= Q].sum() / calls.sum()
    assert frac >= 0.95, frac                       # the likelihood only ever sees whole queues
    assert s.results['eff'] > 1.0


def test_reflective_and_dead_points():
    ndim = 2
    sigma = 0.2

    def lnprob(theta):                              # peak on the wall u0 = 0, NaN model in one corner
        th = np.asarray(theta)
        out = -0.5 * (th[:, 0] / sigma) ** 2 - 0.5 * ((th[:, 1] - 0.5) / sigma) ** 2
        out[(th[:, 0] > 0.9) & (th[:, 1] > 0.9)] = np.nan
        return out
    s = BatchedNestedSampler(lnprob, lambda u: np.asarray(u, dtype=float), ndim, nlive=150, walks=20,
                             reflective=[0], seed=5)
    r = s.run_nested(dlogz=0.05)
    # Z = int_0^1 exp(-x^2/2s^2) dx * int_0^1 exp(-(y-.5)^2/2s^2) dy
    from scipy.stats import norm
    z = (sigma * math.sqrt(2 * math.pi)) ** 2 * (norm.cdf(1 / sigma) - 0.5) * (2 * norm.cdf(0.5 / sigma) - 1)
    assert abs(r['logz'] - math.log(z)) < 4.0 * max(r['logzerr'], 0.08)
    assert np.isfinite(r['logz'])


def test_pool_shim_under_a_dynesty_like_mapper():
    """dynesty with ``pool=`` maps ``evolve_point(args)`` -- a whole random walk with its likelihood calls
    inside -- through ``pool.map`` (its ``Sampler._fill_queue``); only the initial live points go through
    ``pool.map(loglikelihood, points)``.  A stub with that protocol shows what ``BatchedPool`` can and
    cannot batch: the live-point evaluation, not the walks.  The lock-step sampler is the remedy."""
    seen = []

    class Like:
        ndim, parsdict, fitpars_i = 2, {}, ['a', 'b']

        def lnlikefn(self, p):
            seen.append(1)
            return -0.5 * float(np.sum(np.square(p)))

        def lnlike_batch(self, th):
            seen.append(len(th))
            return -0.5 * np.sum(np.square(th), axis=1)

    class Prior:
        def lnpriorfn(self, d):
            return 0.0

        def lnprior_batch(self, th):
            return np.zeros(len(th))
    f = BatchedLnProb(Like(), Prior())
    pool = BatchedPool(queue_size=16)
    rng = np.random.default_rng(0)
    live = rng.random((16, 2))
    # (1) initial live points: pool.map(loglikelihood, points)  -> one batch of 16
    pool.map(f, list(live))
    assert seen == [16]

    # (2) the hot loop: pool.map(evolve_point, argslist); evolve_point calls the likelihood itself
    def evolve_point(args):
        u, loglstar = args
        for _ in range(5):
            up = u + 0.01 * rng.standard_normal(2)
            if f(up) > loglstar:
                u = up
        return u
    del seen[:]
    pool.map(evolve_point, [(u, -10.0) for u in live])
    assert seen == [1] * (16 * 5)                   # every walk step arrived alone
