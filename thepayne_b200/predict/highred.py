"""High-extinction (Av >= 5) analytic extension of the bolometric corrections.

Mirror of ``Payne/predict/highred.py:4-25``: per band ``BC0 - (a1 + b1*Av*(a2 + b2*Rv + c2*Rv**2))``.
The per-band coefficient table of the reference (highred.py:29-169) is data, not code: it ships
as ``thepayne_b200/data/highav_coeffs.txt`` (``filter a1 b1 a2 b2 c2``), written from a reference
checkout by ``tools/export_highav.py``; ``PAYNE_HIAV_TABLE`` / ``table=`` add or override rows.
Bands missing from the table get NaN coefficients, as in the reference (highred.py:14-15) -- with
a warning, because every magnitude of such a band turns NaN as soon as a point has Av >= 5.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

_TABLE_ENV = 'PAYNE_HIAV_TABLE'
TABLE_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'data', 'highav_coeffs.txt')
_cache = {}


def reference_table():
    """{band: (a1, b1, a2, b2, c2)} of the shipped table (all bands of highred.py:29-169)."""
    if 'tab' not in _cache:
        if not os.path.exists(TABLE_PATH):
            raise IOError('%s is missing: run tools/export_highav.py against a reference checkout' % TABLE_PATH)
        _cache['tab'] = _read_table(TABLE_PATH)
    return dict(_cache['tab'])


def _read_table(path):
    out = {}
    with open(path) as f:
        for ln in f:
            p = ln.split()
            if len(p) == 6 and p[0] != 'filter':
                out[p[0]] = tuple(float(v) for v in p[1:])
    return out


class highAv(object):
    def __init__(self, filters, table=None):
        tab = reference_table()
        path = table or os.environ.get(_TABLE_ENV)
        if path:
            tab.update(_read_table(path))
        missing = [ff for ff in filters if ff not in tab]
        if missing:
            warnings.warn('no high-Av coefficients for %s: their magnitudes are NaN for Av >= 5 '
                          '(as in the reference, highred.py:14-15)' % ', '.join(missing))
        self.Avlist = [list(tab[ff]) if ff in tab else [np.nan] * 5 for ff in filters]

    def getAvaprox(self, Av, Rv, pars):
        a1, b1, a2, b2, c2 = pars
        return a1 + b1 * Av * (a2 + b2 * Rv + c2 * Rv ** 2.0)

    def calc(self, BC0, Av, Rv):
        return np.array([b - self.getAvaprox(Av, Rv, p) for p, b in zip(self.Avlist, BC0)])

    @staticmethod
    def export_reference_table(reference_highred_py, dst):
        """Pull the coefficient rows out of a reference checkout's highred.py into ``dst``."""
        rows = []
        with open(reference_highred_py) as f:
            for ln in f:
                p = ln.split()
                if len(p) == 6:
                    try:
                        [float(v) for v in p[1:]]
                    except ValueError:
                        continue
                    rows.append(' '.join(p))
        with open(dst, 'w') as f:
            f.write('filter a1 b1 a2 b2 c2\n' + '\n'.join(rows) + '\n')
        return dst
