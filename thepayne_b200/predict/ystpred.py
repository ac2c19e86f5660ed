"""Legacy spectrum predictor with the reference's names (``Payne/predict/ystpred.py``).

``Net`` (:18-58) is the three-layer leaky-ReLU emulator stored as ``w_array_{0,1,2}`` /
``b_array_{0,1,2}`` / ``x_min`` / ``x_max`` / ``wavelength`` / ``resolution``;
``PayneSpecPredict`` (:60-278) runs the same broadening chain as ``predictspec.PayneSpecPredict``.
``GenMod._initspecnn`` selects this module for ``NNtype='YST1'`` (``genmod.py:18-21``).  The
arithmetic runs in libpayne_b200.so: the leaky-ReLU stack on the CUDA-core fp32 layers, the tail
unchanged.
"""
from __future__ import annotations

import numpy as np

from . import predictspec as _ps


class Net(_ps.ANN):
    """ystpred.py:18-58 (``eval`` -> flux of one label vector; here also batches)."""

    def __init__(self, NNpath, **kwargs):
        super().__init__(nnpath=NNpath, NNtype='YST1', **kwargs)
        # ystpred.py:25-37 attribute names
        self.w_array_0, self.w_array_1, self.w_array_2 = self.model.weights
        self.b_array_0, self.b_array_1, self.b_array_2 = self.model.biases


class PayneSpecPredict(_ps.PayneSpecPredict):
    """ystpred.py:60-278."""

    def __init__(self, nnpath=None, **kwargs):
        self.NN = {}
        self.nnpath = nnpath
        self.NNtype = kwargs.get('NNtype', 'YST1')
        self.Cnnpath = kwargs.get('Cnnpath', None)
        if self.Cnnpath is not None:
            raise NotImplementedError('continuum ANN (Cnnpath) is outside the accelerated path')
        self.anns = Net(nnpath, precision=kwargs.get('precision', 'parity'))
        # networks trained on Teff/1000 are rescaled on load (ystpred.py:76-79)
        if self.anns.xmin[0] < 1000.0:
            m = self.anns.model
            m.xmin = np.array(m.xmin, dtype=np.float64)
            m.xmax = np.array(m.xmax, dtype=np.float64)
            m.xmin[0] *= 1000.0
            m.xmax[0] *= 1000.0
            self.anns.xmin, self.anns.xmax = m.xmin, m.xmax
        self.Canns = None
