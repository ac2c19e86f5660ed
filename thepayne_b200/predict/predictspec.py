"""Spectrum emulator front end with the reference's class and method names.

Mirror of ``Payne/predict/predictspec.py``: ``ANN`` (:29-74) and ``PayneSpecPredict`` (:77-298).
The arithmetic runs in libpayne_b200.so (tcgen05 MLP + fused broadening tail); these classes
only translate the reference's keyword conventions into rows of a parameter matrix.
"""
from __future__ import annotations

import numpy as np

from .. import annio
from ..engine import Engine
from ..synth import SpecNet

speedoflight = 299792.458
_FULL = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R']


def _as_specnet(nnpath, NNtype='LinNet'):
    if isinstance(nnpath, SpecNet):
        return nnpath
    if nnpath is None:
        raise IOError('no spectrum ANN given (the reference default data/ANN/NN.h5 is not shipped)')
    return annio.load_specnet(nnpath, NNtype=NNtype)


class ANN(object):
    """predictspec.py:29-74 -- holds the network and evaluates it, batch-aware."""

    def __init__(self, nnpath=None, **kwargs):
        self.verbose = kwargs.get('verbose', False)
        self.nnpath = nnpath
        self.NNtype = kwargs.get('NNtype', 'LinNet')
        if self.NNtype not in ('LinNet', 'SMLP', 'YST1', 'MultiNet'):
            raise NotImplementedError('NNtype=%r is not accelerated (LinNet, SMLP, YST1 and MultiNet are)' % self.NNtype)
        self.model = _as_specnet(nnpath, self.NNtype)
        if self.model.nntype != self.NNtype:
            raise ValueError('NNtype=%r but the network container holds a %s' % (self.NNtype, self.model.nntype))
        self.inlabels = list(self.model.inlabels)
        self.xmin, self.xmax = self.model.xmin, self.model.xmax
        self.wavelength = self.model.wavelength
        self.resolution = np.array(self.model.resolution, dtype=float)
        self.precision = kwargs.get('precision', 'parity')
        self.continuum = None         # SpecNet of PayneSpecPredict's Cnnpath; attached to every engine made here
        self._engines = {}

    def engine_for(self, outwave=None, npoly=0, inst_sigma=False, lsf=None):
        """Engine whose observed grid is ``outwave`` (None -> the emulator's own grid).
        ``inst_sigma``: the Inst_R column is the sigma-resolution ``getspec`` takes (predictspec.py:255-263),
        not the FWHM resolution ``GenMod.genspec`` multiplies by 2.355 (genmod.py:82-85).
        ``lsf``: dispersion (AA) per pixel of ``outwave`` -- the LSF-vector form of ``inst_R``
        (predictspec.py:265-286)."""
        ow = self.wavelength if outwave is None else np.ascontiguousarray(outwave, dtype=np.float64)
        lsf = None if lsf is None else np.ascontiguousarray(lsf, dtype=np.float64)
        key = (ow.tobytes(), npoly, bool(inst_sigma), None if lsf is None else lsf.tobytes())
        if key not in self._engines:
            if len(self._engines) > 8:
                self._engines.pop(next(iter(self._engines))).close()
            one = np.ones(len(ow))
            fit = _FULL + ['pc_%d' % k for k in range(npoly)]
            self._engines[key] = Engine(spec=self.model, obs_wave=ow, obs_flux=one, obs_eflux=one,
                                        fitpars_i=fit, runbools=(True, False, npoly > 0, False, False),
                                        precision=self.precision)
            if inst_sigma:
                self._engines[key].set('inst_r_is_sigma', 1)
            if self.continuum is not None:
                self._engines[key].attach_continuum(self.continuum)
            if lsf is not None:
                self._engines[key].set_lsf(lsf)
        return self._engines[key]

    def eval(self, x):
        """labels [D_in] or [B, D_in] -> flux [D_out] or [B, D_out] float32 (predictspec.py:61-74)."""
        if isinstance(x, list):
            x = np.asarray(x)
        x = np.asarray(x, dtype=np.float64)
        y = self.engine_for().ann_eval(x.reshape(-1, self.model.D_in))
        return y.cpu().numpy().squeeze()


class PayneSpecPredict(object):
    """predictspec.py:77-298."""

    def __init__(self, nnpath=None, **kwargs):
        self.NN = {}
        self.nnpath = nnpath
        self.NNtype = kwargs.get('NNtype', 'LinNet')
        self.C_NNtype = kwargs.get('C_NNtype', 'LinNet')
        self.anns = ANN(nnpath=nnpath, NNtype=self.NNtype, testing=False, verbose=False,
                        precision=kwargs.get('precision', 'parity'))
        self.Cnnpath = kwargs.get('Cnnpath', None)
        if self.Cnnpath is not None:          # predictspec.py:96-102
            self.Canns = ANN(nnpath=self.Cnnpath, NNtype=self.C_NNtype, testing=False, verbose=False,
                             precision=kwargs.get('precision', 'parity'))
            self.anns.continuum = self.Canns.model
        else:
            self.Canns = None

    def predictspec(self, labels):
        return self.anns.eval(labels)

    def predictcont(self, labels):
        """predictspec.py:122-134."""
        return self.Canns.eval(labels)

    def _labels(self, kwargs):
        """Keyword aliases of predictspec.py:154-198."""
        d = {}
        d['teff'] = kwargs['Teff'] if 'Teff' in kwargs else (10.0 ** kwargs['logt'] if 'logt' in kwargs else 5770.0)
        d['logg'] = kwargs.get('log(g)', kwargs.get('logg', 4.44))
        d['feh'] = kwargs.get('[Fe/H]', kwargs.get('feh', 0.0))
        afe = 0.0
        for k in ['[alpha/Fe]', '[a/Fe]', 'aFe', 'afe']:
            if k in kwargs:
                afe = kwargs[k]
                break
        d['afe'] = afe
        vm = kwargs.get('vmic', np.nan)
        d['vmic'] = vm if np.isfinite(vm) else np.nan
        return d

    def getspec(self, **kwargs):
        """Model spectrum for one label set (predictspec.py:136-294).  ``inst_R``: a float (sigma-resolution)
        or a vector of dispersions in AA, one per pixel of ``outwave`` (of the native grid when ``outwave`` is
        None) -- the LSF case of predictspec.py:265-286."""
        self.inputdict = self._labels(kwargs)
        outwave = kwargs.get('outwave', None)
        rot = kwargs.get('rot_vel', 0.0)
        rad = kwargs.get('rad_vel', 0.0)
        inst = kwargs.get('inst_R', np.nan)
        lsf = None
        if 'inst_R' in kwargs and not isinstance(inst, float):     # predictspec.py:255, 265
            lsf = np.asarray(inst, dtype=np.float64)
        modwave = self.anns.wavelength
        if outwave is None:
            # no resampling of the wavelength axis: the reference returns the (shifted) native grid
            grid = modwave * (1.0 + (rad / speedoflight)) if rad != 0.0 else modwave
        else:
            outwave = np.array(outwave)
            grid = outwave
        if lsf is not None and len(lsf) != len(grid):              # predictspec.py:276-281
            print('Length of LSF vector not equal to input wavelength')
            raise AssertionError
        eng = self.anns.engine_for(grid, inst_sigma=True, lsf=lsf)
        d = self.inputdict
        row = np.array([[d['teff'], d['logg'], d['feh'], d['afe'], rad, rot, d['vmic'],
                         inst if (lsf is None and inst > 0.0) else np.nan]], dtype=np.float64)
        flux, _, _ = eng.model_batch(row, want_mags=False)
        return grid, flux[0].cpu().numpy()

    def getspec_batch(self, labels, rot_vel, rad_vel, inst_R, outwave, vmic=None, lsf=None):
        """Batched extension: arrays of length B -> flux [B, len(outwave)] (CUDA tensor).  ``lsf``: one
        dispersion vector (AA per pixel of ``outwave``) shared by the batch; ``inst_R`` is then ignored."""
        labels = np.asarray(labels, dtype=np.float64)
        B = labels.shape[0]
        th = np.full((B, 8), np.nan)
        th[:, :4] = labels[:, :4]
        th[:, 4], th[:, 5] = rad_vel, rot_vel
        if vmic is not None:
            th[:, 6] = vmic
        if lsf is None:
            th[:, 7] = np.asarray(inst_R, dtype=np.float64)
        flux, _, _ = self.anns.engine_for(outwave, inst_sigma=True, lsf=lsf).model_batch(th, want_mags=False)
        return flux
