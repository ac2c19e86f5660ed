"""Feeding a nested sampler's proposal queue to the GPU in one launch.

The reference registers ``lnprobfn(pars, likeobj, priorobj)`` with dynesty (fitstar.py:309-313,
647-659) and evaluates one vector per call.  dynesty can instead hand a *queue* of proposals to
``pool.map(loglikelihood, points)`` (``NestedSampler(..., pool=P, queue_size=Q)``); ``BatchedPool``
is such a pool-like object: when the mapped callable is (a wrapper of) a ``BatchedLnProb`` the whole
queue becomes one ``likelihood.lnlike_batch`` call, anything else falls back to a plain ``map``.

dynesty is not installed in the build image and the reference does not pin a version, so this adapter
is exercised against a local stand-in sampler only (``tests/test_batching.py``).
"""
from __future__ import annotations

import numpy as np


class BatchedLnProb(object):
    """``lnprobfn`` of fitstar.py:647-659 as a callable object, plus ``batch`` for [B, ndim]."""

    def __init__(self, likeobj, priorobj):
        self.likeobj, self.priorobj = likeobj, priorobj

    def __call__(self, pars, *args):
        lnlike = self.likeobj.lnlikefn(pars)
        if lnlike == -np.inf:
            return -np.inf
        lnprior = self.priorobj.lnpriorfn(self.likeobj.parsdict)
        if lnprior == -np.inf:
            return -np.inf
        return lnprior + lnlike

    def batch(self, theta):
        theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        lnlike = np.asarray(self.likeobj.lnlike_batch(theta), dtype=np.float64)
        lnprior = self.priorobj.lnprior_batch(theta)
        out = lnprior + lnlike
        out[(lnlike == -np.inf) | (lnprior == -np.inf)] = -np.inf
        return out


def _unwrap(func):
    seen = 0
    while seen < 4 and not isinstance(func, BatchedLnProb):
        nxt = getattr(func, 'func', None) or getattr(func, '__wrapped__', None)
        if nxt is None:
            break
        func, seen = nxt, seen + 1
    return func


class BatchedPool(object):
    """Pool-like object for ``dynesty.NestedSampler(pool=..., queue_size=...)``."""

    def __init__(self, queue_size=256):
        self.size = int(queue_size)
        self.batches = 0
        self.points = 0

    def map(self, func, iterable):
        pts = list(iterable)
        target = _unwrap(func)
        if isinstance(target, BatchedLnProb) and len(pts) > 0:
            self.batches += 1
            self.points += len(pts)
            return list(target.batch(np.array(pts, dtype=np.float64)))
        return list(map(func, pts))

    def close(self):
        pass

    def join(self):
        pass
