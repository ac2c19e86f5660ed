"""High-extinction (Av >= 5) analytic extension of the bolometric corrections.

Mirror of ``Payne/predict/highred.py:4-25``: per band ``BC0 - (a1 + b1*Av*(a2 + b2*Rv + c2*Rv**2))``.
The per-band coefficient table of the reference (highred.py:29-169) is data, not code: it is
read from a whitespace table (``filter a1 b1 a2 b2 c2``) so it can be exported from a
reference checkout with ``highAv.export_reference_table`` and dropped next to the ANNs.
Bands missing from the table get NaN coefficients, as in the reference (highred.py:14-15).
"""
from __future__ import annotations

import os

import numpy as np

from ..synth import _HIAV_SYNTH

_TABLE_ENV = 'PAYNE_HIAV_TABLE'


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
        tab = dict(_HIAV_SYNTH)
        path = table or os.environ.get(_TABLE_ENV)
        if path:
            tab.update(_read_table(path))
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
