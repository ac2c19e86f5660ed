"""Static nested sampling whose random walks advance in LOCK-STEP, so that every likelihood
evaluation reaches the GPU as part of one ``lnlike_batch`` call.

Why not a pool shim: the reference builds ``dynesty.NestedSampler`` without a pool
(fitstar.py:309-321) and pulls one dead point per iteration from ``sample()`` (:332-336); with
``pool=`` dynesty maps *proposal evolutions* (``evolve_point``: a whole ``walks``-step random walk
with its likelihood calls inside), not likelihood calls, over the pool -- the likelihood still
arrives one vector at a time.  What batches is the walk itself:

    fill the queue:  Q walkers start from random live points; at each of the ``walks`` steps ALL of
                     them propose (same proposal law as dynesty's ``sample_rwalk``: a uniform draw
                     from the unit n-ball, mapped through the axes of the live points' bounding
                     ellipsoid and multiplied by ``scale``), the Q proposals go through ONE
                     ``lnprob_batch`` call, and each walker moves if its proposal lies above the
                     current likelihood threshold;
    iterate:         the worst live point dies (volume shrinks by 1/nlive on average, evidence and
                     information updated with the trapezoid rule exactly as in dynesty's static
                     sampler); queued proposals are popped until one is still above the -- meanwhile
                     risen -- threshold (a point drawn uniformly inside an earlier, larger contour
                     and found inside the current one is uniform inside the current one);
    adapt:           ``scale`` follows the walkers' acceptance fraction (target 0.5) with dynesty's
                     ``update_rwalk`` rule, applied after every lock-step move.

``sample()`` yields the same 15-tuple as ``dynesty.NestedSampler.sample`` so that the logging loop of
fitstar.py:332-405 carries over; ``add_live_points()`` likewise.  dynesty itself is not installed in
the build image (and unpinned in the reference's setup.py), so this sampler is validated on analytic
evidences (tests/test_nested.py), not against dynesty runs.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import logsumexp


class BatchedNestedSampler(object):
    def __init__(self, lnprob_batch, prior_transform_batch, ndim, nlive=125, walks=25, queue_size=None,
                 facc=0.5, reflective=(), seed=None, max_extra_walks=0):
        """``lnprob_batch(theta[B, ndim]) -> lnP[B]`` (``BatchedLnProb.batch``) and
        ``prior_transform_batch(U[B, ndim]) -> theta[B, ndim]`` (``prior.priortrans_batch``)."""
        self.lnprob_batch = lnprob_batch
        self.ptform = prior_transform_batch
        self.ndim, self.nlive, self.walks = int(ndim), int(nlive), int(walks)
        self.Q = int(queue_size) if queue_size else self.nlive
        self.facc = float(facc)
        self.reflective = np.zeros(self.ndim, dtype=bool)
        self.reflective[list(reflective)] = True
        self.rng = np.random.default_rng(seed)
        self.max_extra = int(max_extra_walks)
        self.scale = 1.0
        self.f_inside = 0.5           # fraction of proposals inside the unit cube (estimate for the next fill)
        self.batch_sizes = []          # size of every lnprob_batch call (the batching evidence)
        self.ncall = 0
        self.it = 0
        self.queue = []
        # initial live points: one batch
        self.live_u = self.rng.random((self.nlive, self.ndim))
        self.live_v = np.asarray(self.ptform(self.live_u), dtype=np.float64)
        self.live_logl = self._eval(self.live_v)
        self.live_it = np.zeros(self.nlive, dtype=int)
        # run state (dynesty's static sampler bookkeeping)
        self.dlv = math.log((self.nlive + 1.0) / self.nlive)
        self.logvol, self.logz, self.logzvar, self.h, self.loglstar = 0.0, -1e300, 0.0, 0.0, -1e300
        self.saved = {k: [] for k in ['u', 'v', 'logl', 'logvol', 'logwt', 'logz', 'logzvar', 'h', 'nc', 'it']}
        self.added_live = False

    # ------------------------------------------------------------------ likelihood calls
    def _eval(self, v):
        out = np.asarray(self.lnprob_batch(np.ascontiguousarray(v)), dtype=np.float64).copy()
        out[~(out == out)] = -np.inf                   # a NaN model never beats a threshold
        self.batch_sizes.append(len(v))
        self.ncall += len(v)
        return out

    # ------------------------------------------------------------------ proposals
    def _axes(self):
        """Axes of the ellipsoid bounding the live points in the unit cube (covariance x (n+2))."""
        cov = np.cov(self.live_u, rowvar=False).reshape(self.ndim, self.ndim) * (self.ndim + 2.0)
        cov += 1e-14 * np.eye(self.ndim) * max(1.0, np.trace(cov))
        try:
            return np.linalg.cholesky(cov)
        except np.linalg.LinAlgError:
            w, V = np.linalg.eigh(cov)
            return V * np.sqrt(np.maximum(w, 1e-16))

    def _propose(self, u, axes):
        """One rwalk proposal per row of u and the mask of those inside the unit cube.  A proposal
        outside the cube is a REJECTED STEP of that walker (it stays put and takes no likelihood call).
        Redrawing it instead -- tempting, because it keeps the batch full -- makes the proposal law depend
        on the distance to the walls and biased ln Z by +0.3 on a 4-d Gaussian test."""
        n = len(u)
        dr = self.rng.standard_normal((n, self.ndim))
        dr /= np.linalg.norm(dr, axis=1, keepdims=True)
        dr *= self.rng.random((n, 1)) ** (1.0 / self.ndim)
        p = u + self.scale * (dr @ axes.T)
        if self.reflective.any():                       # fold back at 0 and 1 (dynesty's reflective walls)
            r = self.reflective
            q = np.mod(p[:, r], 2.0)
            p[:, r] = np.where(q > 1.0, 2.0 - q, q)
        return p, np.all((p > 0.0) & (p < 1.0), axis=1)

    def _fill_queue(self, loglstar):
        """W walkers x ``walks`` lock-step Metropolis steps.  W >= Q is chosen so that the proposals that
        fall inside the cube -- the ones that are evaluated -- number at least Q per step (the inside
        fraction of the previous fill is the estimate; it tends to 1 as the contours leave the walls)."""
        axes = self._axes()
        cand = np.flatnonzero(self.live_logl > loglstar)
        if len(cand) == 0:
            cand = np.arange(self.nlive)
        W = min(int(math.ceil(1.08 * self.Q / max(self.f_inside, 0.34))), 3 * self.Q)
        start = self.rng.choice(cand, size=W, replace=True)
        u, v, logl = self.live_u[start].copy(), self.live_v[start].copy(), self.live_logl[start].copy()
        nc = np.zeros(W, dtype=int)
        nacc = np.zeros(W, dtype=int)
        nrej = np.zeros(W, dtype=int)
        active = np.arange(W)
        step, f_min = 0, 1.0
        while len(active):
            pu, inside = self._propose(u[active], axes)
            if step < self.walks:
                f_min = min(f_min, float(inside.mean()))
            nrej[active[~inside]] += 1
            ev = active[inside]
            if len(ev):
                pu = pu[inside]
                pv = np.asarray(self.ptform(pu), dtype=np.float64)
                pl = self._eval(pv)
                nc[ev] += 1
                ok = pl > loglstar
                idx = ev[ok]
                u[idx], v[idx], logl[idx] = pu[ok], pv[ok], pl[ok]
                nacc[idx] += 1
                nrej[ev[~ok]] += 1
            step += 1
            # step size follows the acceptance fraction of THIS lock-step move (W outcomes at once: as much
            # evidence as dynesty's per-proposal update_rwalk collects over W iterations)
            facc = float(ok.sum()) / max(len(active), 1) if len(ev) else 0.0
            norm = max(self.facc, 1.0 - self.facc) * self.ndim
            self.scale = min(self.scale * math.exp((facc - self.facc) / norm), math.sqrt(self.ndim))
            if step >= self.walks:
                # dynesty lets a walk that has not moved yet go on until it does; in lock-step those
                # stragglers would arrive in calls of a handful of points, so by default
                # (max_extra_walks = 0) a walker that never moved simply yields no proposal
                active = active[nacc[active] == 0]
                if step >= self.walks * (1 + self.max_extra):
                    break
        self.f_inside = f_min
        for i in range(W):
            if nacc[i] > 0:
                self.queue.append((u[i], v[i], logl[i], int(nc[i])))
            else:
                self.queue.append((None, None, -np.inf, int(nc[i])))      # never moved: only its calls count

    def _new_point(self, loglstar):
        nc = 0
        while True:
            if not self.queue:
                self._fill_queue(loglstar)
            u, v, logl, c = self.queue.pop()
            nc += c
            if u is not None and logl > loglstar:
                return u, v, logl, nc

    # ------------------------------------------------------------------ evidence bookkeeping
    def _accumulate(self, loglstar_new, logvol):
        logdvol = logsumexp(a=[logvol + self.dlv_cur, logvol], b=[0.5, -0.5])
        logwt = np.logaddexp(loglstar_new, self.loglstar) + logdvol
        logz_new = np.logaddexp(self.logz, logwt)
        lzterm = (math.exp(self.loglstar - logz_new) * self.loglstar +
                  math.exp(loglstar_new - logz_new) * loglstar_new) if self.loglstar > -1e299 else \
            math.exp(loglstar_new - logz_new) * loglstar_new
        h_new = math.exp(logdvol) * lzterm + math.exp(self.logz - logz_new) * (self.h + self.logz) - logz_new
        dh = h_new - self.h
        self.h, self.logz = h_new, logz_new
        self.logzvar += dh * self.dlv_cur              # Var(ln Z) ~ H / nlive (Skilling 2006)
        self.loglstar = loglstar_new
        return logwt

    def _save(self, u, v, logl, logvol, logwt, nc, it):
        s = self.saved
        s['u'].append(np.array(u)); s['v'].append(np.array(v)); s['logl'].append(logl); s['logvol'].append(logvol)
        s['logwt'].append(logwt); s['logz'].append(self.logz); s['logzvar'].append(self.logzvar)
        s['h'].append(self.h); s['nc'].append(nc); s['it'].append(it)

    # ------------------------------------------------------------------ the two generators
    def sample(self, dlogz=0.01, maxiter=None, maxcall=None):
        maxiter = np.inf if maxiter is None else maxiter
        maxcall = np.inf if maxcall is None else maxcall
        while True:
            logz_remain = np.max(self.live_logl) + self.logvol
            delta_logz = np.logaddexp(self.logz, logz_remain) - self.logz
            if self.it > 0 and delta_logz < dlogz:
                break
            if self.it >= maxiter or self.ncall >= maxcall:
                break
            worst = int(np.argmin(self.live_logl))
            worst_it = int(self.live_it[worst])
            ustar, vstar = self.live_u[worst].copy(), self.live_v[worst].copy()
            loglstar_new = float(self.live_logl[worst])
            if not np.isfinite(loglstar_new) and loglstar_new < 0:
                loglstar_new = -1e300                  # dead-on-arrival points carry no weight
            self.logvol -= self.dlv
            self.dlv_cur = self.dlv
            logwt = self._accumulate(loglstar_new, self.logvol)
            u, v, logl, nc = self._new_point(loglstar_new)
            self._save(ustar, vstar, loglstar_new, self.logvol, logwt, nc, worst_it)
            self.live_u[worst], self.live_v[worst], self.live_logl[worst] = u, v, logl
            self.live_it[worst] = self.it + 1
            self.it += 1
            eff = 100.0 * self.it / self.ncall
            logz_remain = np.max(self.live_logl) + self.logvol
            delta_logz = np.logaddexp(self.logz, logz_remain) - self.logz
            yield (worst, ustar, vstar, loglstar_new, self.logvol, logwt, self.logz, self.logzvar, self.h, nc,
                   worst_it, 0, self.it, eff, delta_logz)

    def add_live_points(self):
        """The remaining live points in order of likelihood, each with the expected volume it bounds
        (dynesty's ``add_live_points``)."""
        if self.added_live:
            return
        self.added_live = True
        order = np.argsort(self.live_logl)
        logvol0 = self.logvol
        for i, idx in enumerate(order):
            logvol = logvol0 + math.log(1.0 - (i + 1.0) / (self.nlive + 1.0))
            self.dlv_cur = self.logvol - logvol
            self.logvol = logvol
            ll = float(self.live_logl[idx])
            if not np.isfinite(ll) and ll < 0:
                ll = -1e300
            logwt = self._accumulate(ll, logvol)
            self._save(self.live_u[idx], self.live_v[idx], ll, logvol, logwt, 1, int(self.live_it[idx]))
            eff = 100.0 * (self.it + i) / self.ncall
            yield (int(idx), self.live_u[idx].copy(), self.live_v[idx].copy(), ll, logvol, logwt, self.logz,
                   self.logzvar, self.h, 1, int(self.live_it[idx]), 0, self.it, eff, 0.0)

    def run_nested(self, dlogz=0.01, maxiter=None, maxcall=None):
        for _ in self.sample(dlogz=dlogz, maxiter=maxiter, maxcall=maxcall):
            pass
        for _ in self.add_live_points():
            pass
        return self.results

    # ------------------------------------------------------------------ results
    @property
    def results(self):
        s = self.saved
        logwt = np.array(s['logwt'])
        w = np.exp(logwt - self.logz)
        return {'samples': np.array(s['v']), 'samples_u': np.array(s['u']), 'logl': np.array(s['logl']),
                'logvol': np.array(s['logvol']), 'logwt': logwt, 'weights': w / w.sum(), 'logz': self.logz,
                'logzerr': math.sqrt(max(self.logzvar, 0.0)), 'h': self.h, 'ncall': self.ncall, 'niter': self.it,
                'eff': 100.0 * self.it / max(self.ncall, 1), 'batch_sizes': np.array(self.batch_sizes)}

    def posterior_mean_std(self):
        r = self.results
        m = r['weights'] @ r['samples']
        sd = np.sqrt(np.maximum(r['weights'] @ (r['samples'] - m) ** 2, 0.0))
        return m, sd
