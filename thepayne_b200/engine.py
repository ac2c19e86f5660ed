"""Batched likelihood engine: one CUDA context per (emulators, observation, parameter layout).

Thin host-side wrapper over the C ABI (include/payne_b200.h).  PyTorch is used only as the
device-memory / stream provider; the arithmetic all happens inside libpayne_b200.so.

The parameter-vector convention is the reference's (Payne/fitting/likelihood.py:35-72): a
row of ``theta`` holds the free parameters in ``fitpars_i`` order, fixed ones come from
``fixedpars``, and spectral parameters that are neither are NaN.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import PAR_INDEX, PREC


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _specnet_struct(spec, keep):
    """PayneSpecNet (include/payne_b200.h) for a ``synth.SpecNet``-like container; the numpy buffers the
    struct points into are appended to ``keep``."""
    sp = _lib.PayneSpecNet()
    W = [_f32(w) for w in spec.weights]
    b = [_f32(v) for v in spec.biases]
    keep += W + b
    nl = len(W)
    nntype = getattr(spec, 'nntype', 'LinNet')
    if (nl, nntype) not in [(6, 'LinNet'), (4, 'SMLP'), (3, 'YST1'), (4, 'MultiNet')]:
        raise ValueError('unsupported emulator: %d layers of type %s' % (nl, nntype))
    sp.D_in, sp.H1, sp.D_out = W[0].shape[-1], W[0].shape[-2], W[-1].shape[0]
    if nntype == 'MultiNet':   # trainspec_multi.py:29-36, chunk nets stacked along axis 0
        sp.H2 = sp.H3 = sp.H1
        sp.n_groups, sp.group_size = int(W[0].shape[0]), int(spec.chunk)
    elif nl == 6:        # NNmodels.py:147-152
        sp.H2, sp.H3 = W[3].shape[0], W[4].shape[0]
    elif nl == 4:        # NNmodels.py:99-107
        sp.H2, sp.H3 = W[1].shape[0], W[2].shape[0]
    else:                # ystpred.py:25-30
        sp.H2, sp.H3 = W[1].shape[0], 0
    sp.n_layers = nl
    sp.activation = 0 if nntype in ('LinNet', 'MultiNet') else 1
    sp.label_fp32_cast = 0 if nntype == 'YST1' else 1      # ystpred.py:47-50 stays in float64
    for k in range(nl):
        sp.W[k], sp.b[k] = _pf(W[k]), _pf(b[k])
    xmin, xmax, wave = _f64(spec.xmin), _f64(spec.xmax), _f64(spec.wavelength)
    keep += [xmin, xmax, wave]
    sp.xmin, sp.xmax, sp.wavelength = _pd(xmin), _pd(xmax), _pd(wave)
    sp.resolution = float(spec.resolution)
    sp.encode_offset = float(getattr(spec, 'encode_offset', 0.5))
    return sp


class Engine:
    def __init__(self, spec=None, phot=None, obs_wave=None, obs_flux=None, obs_eflux=None,
                 obs_phot=None, fitpars_i=(), fixedpars=None, runbools=(True, False, False, False, False),
                 precision='parity', device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.PayneError('no CUDA device: thepayne_b200 has no CPU fallback')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.fitpars_i = list(fitpars_i)
        self.ndim = len(self.fitpars_i)
        self.fixedpars = dict(fixedpars or {})
        spec_bool, phot_bool, modpoly_bool, photscale_bool = [bool(b) for b in runbools[:4]]
        keep = []   # numpy buffers referenced by the structs until create returns

        lay = _lib.PayneLayout()
        lay.ndim = self.ndim
        for name, k in PAR_INDEX.items():
            lay.col[k] = self.fitpars_i.index(name) if name in self.fitpars_i else -1
            lay.fixed[k] = float(self.fixedpars[name]) if name in self.fixedpars else float('nan')
        pcs = [i for i, p in enumerate(self.fitpars_i) if 'pc' in p]   # likelihood.py:56-57
        if modpoly_bool and len(pcs) > _lib.MAX_POLY:
            raise ValueError('at most %d continuum coefficients' % _lib.MAX_POLY)
        lay.n_poly = len(pcs) if modpoly_bool else 0
        for i, c in enumerate(pcs[:_lib.MAX_POLY]):
            lay.poly_col[i] = c
        lay.spec_bool, lay.phot_bool = int(spec_bool), int(phot_bool)
        lay.modpoly_bool, lay.photscale_bool = int(modpoly_bool), int(photscale_bool)
        lay.precision = PREC[precision] if isinstance(precision, str) else int(precision)

        sp = None
        if spec_bool:
            if spec is None:
                raise ValueError('spec_bool set but no spectrum emulator given')
            sp = _specnet_struct(spec, keep)
            self.D_in, self.D_out = int(sp.D_in), int(sp.D_out)
        ob = _lib.PayneObs()
        self.n_obs = 0
        if spec_bool:
            ow, of, oe = _f64(obs_wave), _f64(obs_flux), _f64(obs_eflux)
            if not (len(ow) == len(of) == len(oe)):
                raise ValueError('observed wave/flux/eflux lengths differ')
            keep += [ow, of, oe]
            ob.n_obs, ob.wave, ob.flux, ob.eflux = len(ow), _pd(ow), _pd(of), _pd(oe)
            self.n_obs = len(ow)
        ph = None
        self.bands = []
        if phot_bool:
            if phot is None or obs_phot is None:
                raise ValueError('phot_bool set but no photometry emulator / observation given')
            self.bands = list(phot.bands)
            ph = _lib.PaynePhotNet()
            arrs = [_f32(phot.w1), _f32(phot.b1), _f32(phot.w2), _f32(phot.b2),
                    _f32(np.asarray(phot.w3).reshape(len(self.bands), -1)), _f32(np.asarray(phot.b3).reshape(-1))]
            keep += arrs
            ph.nb, ph.H = arrs[0].shape[0], arrs[0].shape[1]
            ph.w1, ph.b1, ph.w2, ph.b2, ph.w3, ph.b3 = [_pf(a) for a in arrs]
            pxmin, pxmax, hiav = _f64(phot.xmin), _f64(phot.xmax), _f64(phot.hiav)
            keep += [pxmin, pxmax, hiav]
            ph.xmin, ph.xmax, ph.hiav = _pd(pxmin), _pd(pxmax), _pd(hiav)
            mag = _f64([obs_phot[b][0] for b in self.bands])
            err = _f64([obs_phot[b][1] for b in self.bands])
            keep += [mag, err]
            ob.nb, ob.phot_mag, ob.phot_err = len(self.bands), _pd(mag), _pd(err)
        self.nb = len(self.bands)

        ctx = C.c_void_p()
        _lib.check(self.lib.payne_ctx_create(C.byref(sp) if sp is not None else None,
                                              C.byref(ph) if ph is not None else None,
                                              C.byref(ob), C.byref(lay), self.device, C.byref(ctx)))
        self._ctx = ctx
        del keep

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, '_ctx', None):
            self.lib.payne_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, key):
        return int(self.lib.payne_ctx_query(self._ctx, key.encode()))

    def set(self, key, value):
        if key == 'precision' and isinstance(value, str):
            value = PREC[value]
        _lib.check(self.lib.payne_ctx_set(self._ctx, key.encode(), int(value)))

    def attach_continuum(self, cont):
        """Continuum emulator of ``PayneSpecPredict(Cnnpath=...)`` (predictspec.py:96-102): every model
        spectrum is multiplied by the normalised F_lambda continuum interpolated onto the emulator grid
        (predictspec.py:208-226)."""
        keep = []
        sp = _specnet_struct(cont, keep)
        _lib.check(self.lib.payne_ctx_attach_continuum(self._ctx, C.byref(sp)))

    def set_lsf(self, lsf):
        """LSF vector: the dispersion (AA) at every observed pixel (predictspec.py:265-286); None switches
        back to the scalar ``Inst_R`` stage."""
        if lsf is None:
            _lib.check(self.lib.payne_ctx_set_lsf(self._ctx, None, 0))
            return
        v = _f64(lsf)
        if v.ndim != 1 or len(v) != self.n_obs:
            raise ValueError('the LSF vector needs one dispersion per observed pixel (%d)' % self.n_obs)
        _lib.check(self.lib.payne_ctx_set_lsf(self._ctx, v.ctypes.data, len(v)))

    def last_ms(self, which):
        return float(self.lib.payne_ctx_last_ms(self._ctx, {'mlp': 0, 'tail': 1, 'phot': 2}[which]))

    # ------------------------------------------------------------------ compute
    def _theta_dev(self, theta):
        if not (theta.is_cuda and theta.dtype == torch.float64 and theta.dim() == 2):
            raise ValueError('theta must be a 2-D float64 CUDA tensor')
        if theta.shape[1] != self.ndim:
            raise ValueError('theta has %d columns, layout expects %d' % (theta.shape[1], self.ndim))
        return theta.contiguous()

    def lnlike_batch(self, theta):
        """theta: [B, ndim] float64.  CUDA tensor -> CUDA tensor [B] (stream-ordered, no sync);
        numpy / CPU tensor -> numpy [B] through the host-buffer entry point."""
        if isinstance(theta, torch.Tensor) and theta.is_cuda:
            th = self._theta_dev(theta)
            out = torch.empty(th.shape[0], dtype=torch.float64, device=th.device)
            st = torch.cuda.current_stream(th.device).cuda_stream
            _lib.check(self.lib.payne_lnlike_batch(self._ctx, th.data_ptr(), th.shape[0], th.shape[1],
                                                   out.data_ptr(), st))
            return out
        th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64))
        if th.ndim != 2 or th.shape[1] != self.ndim:
            raise ValueError('theta must be [B, %d]' % self.ndim)
        out = np.empty(th.shape[0], dtype=np.float64)
        _lib.check(self.lib.payne_lnlike_batch_host(self._ctx, th.ctypes.data, th.shape[0], th.shape[1],
                                                    out.ctypes.data))
        return out

    def model_batch(self, theta, want_flux=True, want_mags=True):
        """Model spectra [B, n_obs], magnitudes [B, nb] and lnL [B] as CUDA float64 tensors."""
        dev = torch.device('cuda', self.device)
        if not isinstance(theta, torch.Tensor):
            theta = torch.from_numpy(np.ascontiguousarray(np.asarray(theta, dtype=np.float64)))
        th = self._theta_dev(theta.to(dev))
        B = th.shape[0]
        flux = torch.empty((B, self.n_obs), dtype=torch.float64, device=dev) if (want_flux and self.n_obs) else None
        mags = torch.empty((B, self.nb), dtype=torch.float64, device=dev) if (want_mags and self.nb) else None
        lnl = torch.empty(B, dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.lib.payne_model_batch(
            self._ctx, th.data_ptr(), B, th.shape[1],
            flux.data_ptr() if flux is not None else None,
            mags.data_ptr() if mags is not None else None, lnl.data_ptr(), st))
        return flux, mags, lnl

    def ann_eval(self, x):
        """Emulator forward pass: labels [B, D_in] -> flux [B, D_out] float32 (CUDA tensor)."""
        dev = torch.device('cuda', self.device)
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64))))
        x = x.to(dev, dtype=torch.float64).contiguous()
        if x.shape[1] != self.D_in:
            raise ValueError('labels have %d columns, emulator expects %d' % (x.shape[1], self.D_in))
        ldy = (self.D_out + 3) // 4 * 4
        y = torch.empty((x.shape[0], ldy), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(self.lib.payne_ann_eval(self._ctx, x.data_ptr(), x.shape[0], y.data_ptr(), ldy, st))
        return y[:, :self.D_out]


def engine_from_config(cfg, precision='parity', device=None):
    """Engine for a ``thepayne_b200.synth.SynthConfig``."""
    return Engine(spec=cfg.spec if cfg.runbools[0] else None, phot=cfg.phot if cfg.runbools[1] else None,
                  obs_wave=cfg.obs_wave, obs_flux=cfg.obs_flux, obs_eflux=cfg.obs_eflux,
                  obs_phot=cfg.obs_phot, fitpars_i=cfg.fitpars_i, fixedpars=cfg.fixedpars,
                  runbools=cfg.runbools, precision=precision, device=device)
