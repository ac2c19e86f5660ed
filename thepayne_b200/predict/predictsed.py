"""Photometry emulator front end (mirror of ``Payne/predict/predictsed.py:59-103``)."""
from __future__ import annotations

import numpy as np

from .. import annio
from ..engine import Engine
from ..synth import PhotNet
from .highred import highAv


class FastPayneSEDPredict(object):
    def __init__(self, usebands=None, nnpath=None, **kwargs):
        if isinstance(nnpath, PhotNet):
            self.net = nnpath
            usebands = list(nnpath.bands) if usebands is None else list(usebands)
            if usebands != list(nnpath.bands):
                idx = [nnpath.bands.index(b) for b in usebands]
                n = nnpath
                self.net = PhotNet(usebands, n.w1[idx], n.b1[idx], n.w2[idx], n.b2[idx], n.w3[idx], n.b3[idx],
                                   n.xmin, n.xmax, n.hiav[idx])
        else:
            if usebands is None:
                raise IOError('usebands must be given when the ANNs are read from disk')
            self.net = annio.load_photnet(nnpath, list(usebands))
        self.filternames = list(usebands)
        self.HiAv = highAv(self.filternames)
        self.anns = self.net
        self._eng = {}
        self.precision = kwargs.get('precision', 'parity')

    def _engine(self, scaled):
        if scaled not in self._eng:
            fit = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]'] + (['log(A)'] if scaled else ['log(R)', 'Dist']) + ['Av']
            dummy = {b: [0.0, 1.0] for b in self.filternames}
            self._eng[scaled] = Engine(phot=self.net, obs_phot=dummy, fitpars_i=fit,
                                       runbools=(False, True, False, scaled, False), precision=self.precision)
        return self._eng[scaled]

    def sed(self, logt=None, logg=None, feh=None, afe=None, logl=None, av=0.0, rv=3.1,
            dist=None, logA=None, band_indices=slice(None)):
        """Apparent magnitudes for one star (predictsed.py:75-103).  ``rv`` other than 3.1 is not
        reachable from the likelihood (likelihood.py:103-106) and is not supported here."""
        if rv is not None and rv != 3.1:
            raise NotImplementedError('Rv != 3.1 is outside the accelerated path')
        teff = 10.0 ** logt
        if (logl is not None) and (dist is not None):
            logR = 0.5 * (logl - 4.0 * (logt - np.log10(5770.0)))
            row = np.array([[teff, logg, feh, afe, logR, dist, av]])
            m = self._engine(False).model_batch(row, want_flux=False)[1]
        elif logA is not None:
            row = np.array([[teff, logg, feh, afe, logA, av]])
            m = self._engine(True).model_batch(row, want_flux=False)[1]
        else:
            raise IOError('cannot understand input pars into sed function')
        m = m[0].cpu().numpy()
        try:
            return m[band_indices]
        except IndexError:
            return [m]
