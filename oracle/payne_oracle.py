"""CPU ORACLE for the batched-likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy (+ torch CPU for the fp32 Linear/sigmoid the
reference itself calls), the algorithm of pacargile/ThePayne's likelihood path:

    likelihood.lnlikefn -> GenMod.genspec -> PayneSpecPredict.getspec
        -> LinNet.forward -> smoothspec('vsini') -> Doppler -> smoothspec('R')
        -> polycalc -> chi2 (+ FastPayneSEDPredict.sed -> chi2)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package ``thepayne_b200`` never
does (it fails loudly when its CUDA library is missing).

PARITY PIN: the reference has no tests or golden vectors of its own,
so this oracle is pinned against outputs of the *unmodified reference code* executed
in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``), checked
by ``tests/test_oracle_golden.py``.

Every function cites the reference lines it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import numpy as np
import torch
from numpy.polynomial.chebyshev import chebval
from scipy.special import j1

CKMS = 2.998e5                    # Payne/utils/smoothing.py:16
SPEEDOFLIGHT = 299792.458         # scipy.constants.c/1000, Payne/predict/predictspec.py:12
FWHM_FACTOR_FIT = 2.355           # Payne/fitting/genmod.py:83
LOG10_TSUN = np.log10(5770.0)


# --------------------------------------------------------------------------- emulator
class TorchLinNet:
    """fp32 forward pass: ``NNmodels.py:154-168`` (LinNet) driven the way
    ``predictspec.py:61-74`` (ANN.eval) drives it."""

    def __init__(self, spec, ideal=False):
        """``ideal=True`` evaluates the same network in float64 (exact-arithmetic yardstick used
        by the tests to measure the reference's own fp32 round-off; not a reference code path)."""
        self.spec = spec
        self.ideal = ideal
        dt = torch.float64 if ideal else torch.float32
        self.W = [torch.from_numpy(np.ascontiguousarray(w)).to(dt) for w in spec.weights]
        self.b = [torch.from_numpy(np.ascontiguousarray(b)).to(dt) for b in spec.biases]

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        # ANN.eval casts to a FloatTensor first (predictspec.py:70) ...
        x32 = torch.from_numpy(x).type(torch.FloatTensor)
        # ... LinNet.encode then works in numpy: fp32 x against fp64 xmin/xmax gives an
        # fp64 result that is cast back to fp32 (NNmodels.py:164-168)
        enc = (x32.numpy() - self.spec.xmin) / (self.spec.xmax - self.spec.xmin) \
            - self.spec.encode_offset
        h = torch.from_numpy(enc).type(torch.FloatTensor)
        if self.ideal:
            h = h.double()
        with torch.no_grad():
            for k in range(5):
                h = torch.sigmoid(torch.nn.functional.linear(h, self.W[k], self.b[k]))
            y = torch.nn.functional.linear(h, self.W[5], self.b[5])
        return y.numpy()          # [B, D_out] float32


class LegacyNet:
    """The two leaky-ReLU emulators the reference still accepts.

    ``SMLP`` (``NNmodels.py:92-115``, driven by ``ANN.eval`` ``predictspec.py:61-74``): torch fp32,
    ``Linear, LeakyReLU`` x3 + ``Linear``; encode in numpy on the FloatTensor's values.
    ``YST1`` (``predict/ystpred.py:47-58``): numpy all the way -- float64 labels, the stored weight
    arrays as they are, ``z*(z>0) + 0.01*z*(z<0)``, three ``einsum`` layers."""

    def __init__(self, spec, ideal=False):
        """``ideal=True``: the SMLP evaluated in float64 (yardstick for the reference's own fp32 round-off;
        YST1 already is float64 in the reference)."""
        self.spec = spec
        self.ideal = ideal

    def __call__(self, x):
        sp = self.spec
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        if sp.nntype == 'SMLP' and self.ideal:
            x32 = torch.from_numpy(x).type(torch.FloatTensor)
            enc = (x32.numpy() - sp.xmin) / (sp.xmax - sp.xmin) - 0.5
            h = torch.from_numpy(enc).type(torch.FloatTensor).double()
            with torch.no_grad():
                for k in range(3):
                    h = torch.nn.functional.leaky_relu(torch.nn.functional.linear(
                        h, torch.from_numpy(sp.weights[k]).double(), torch.from_numpy(sp.biases[k]).double()))
                y = torch.nn.functional.linear(h, torch.from_numpy(sp.weights[3]).double(),
                                               torch.from_numpy(sp.biases[3]).double())
            return y.numpy()
        if sp.nntype == 'SMLP':
            x32 = torch.from_numpy(x).type(torch.FloatTensor)
            enc = (x32.numpy() - sp.xmin) / (sp.xmax - sp.xmin) - 0.5
            h = torch.from_numpy(enc).type(torch.FloatTensor)
            W = [torch.from_numpy(np.ascontiguousarray(w)) for w in sp.weights]
            b = [torch.from_numpy(np.ascontiguousarray(v)) for v in sp.biases]
            with torch.no_grad():
                for k in range(3):
                    h = torch.nn.functional.leaky_relu(torch.nn.functional.linear(h, W[k], b[k]))
                y = torch.nn.functional.linear(h, W[3], b[3])
            return y.numpy()
        lrelu = lambda z: z * (z > 0) + 0.01 * z * (z < 0)
        out = []
        for row in x:
            xi = (row - sp.xmin) / (sp.xmax - sp.xmin) - 0.5
            inside = np.einsum('ij,j->i', sp.weights[0], xi) + sp.biases[0]
            outside = np.einsum('ij,j->i', sp.weights[1], lrelu(inside)) + sp.biases[1]
            out.append(np.einsum('ij,j->i', sp.weights[2], lrelu(outside)) + sp.biases[2])
        return np.array(out)      # [B, D_out] float64


class MultiChunkNet:
    """Multi-chunk emulator: one ``Net(D_in, H, P)`` per block of pixels
    (``Payne/train/old/trainspec_multi.py:29-52``): ``encode`` = ``(x - xmin)/(xmax - xmin)`` in numpy on the
    FloatTensor's values, no -0.5 (``:56-67``), then ``sigmoid(lin1), sigmoid(lin2), sigmoid(lin3), lin4``
    in torch fp32; the chunks' outputs side by side make the spectrum."""

    def __init__(self, spec, ideal=False):
        self.spec, self.ideal = spec, ideal

    def __call__(self, x):
        sp = self.spec
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        dt = torch.float64 if self.ideal else torch.float32
        x32 = torch.from_numpy(x).type(torch.FloatTensor)
        enc = (x32.numpy() - sp.xmin) / (sp.xmax - sp.xmin)
        h0 = torch.from_numpy(enc).type(torch.FloatTensor).to(dt)
        out = []
        lin = torch.nn.functional.linear
        with torch.no_grad():
            for g in range(sp.n_groups):
                lo, hi = g * sp.chunk, min((g + 1) * sp.chunk, sp.D_out)
                W = [torch.from_numpy(np.ascontiguousarray(sp.weights[k][g])).to(dt) for k in range(3)]
                b = [torch.from_numpy(np.ascontiguousarray(sp.biases[k][g])).to(dt) for k in range(3)]
                W4 = torch.from_numpy(np.ascontiguousarray(sp.weights[3][lo:hi])).to(dt)
                b4 = torch.from_numpy(np.ascontiguousarray(sp.biases[3][lo:hi])).to(dt)
                h = h0
                for k in range(3):
                    h = torch.sigmoid(lin(h, W[k], b[k]))
                out.append(lin(h, W4, b4).numpy())
        return np.concatenate(out, axis=1)


def make_net(spec, ideal=False):
    if spec.nntype == 'MultiNet':
        return MultiChunkNet(spec, ideal=ideal)
    return TorchLinNet(spec, ideal=ideal) if spec.nntype == 'LinNet' else LegacyNet(spec, ideal=ideal)


# --------------------------------------------------------------------------- smoothing
def _resample_pow2(w, s):
    """``smoothing.py:649-668`` resample_wave (log branch)."""
    nnew = int(2.0 ** np.ceil(np.log2(len(w))))
    lnlam = np.linspace(np.log(w.min()), np.log(w.max()), nnew)
    wn = np.exp(lnlam)
    return wn, np.interp(wn, w, s)


def _mask(wave, width, outwave):
    """``smoothing.py:631-647`` mask_wave, non-linear branch, nsigma_pad=20."""
    if outwave is not None:
        wlim = np.array([outwave.min(), outwave.max()])
    else:
        wlim = np.squeeze(np.array([0, np.inf]))
    wlim = wlim * (1 + 20.0 / width * np.array([-1, 1]))
    return (wave > wlim[0]) & (wave < wlim[1])


def smooth_vsini(wave, spec, vsini):
    """``smoothspec(type='vsini', outwave=None, inres=0)``: ``smoothing.py:93-100,
    132-143, 293-314, 610-629``."""
    with np.errstate(divide='ignore', invalid='ignore'):
        width = CKMS / vsini
        m = _mask(wave, width, None)
    w = wave[m]
    s = np.nan_to_num(spec[m], nan=1.0)
    sigma = np.sqrt(vsini ** 2 - 0.0 ** 2)
    wn, sn = _resample_pow2(w, s)
    dv = CKMS * np.median(np.diff(np.log(wn)))
    ss = np.fft.rfftfreq(len(sn), d=dv)
    ss[0] = 0.01
    ub = 2.0 * np.pi * sigma * ss
    sb = j1(ub) / ub - 3 * np.cos(ub) / (2 * ub ** 2) + 3.0 * np.sin(ub) / (2 * ub ** 3)
    sb[0] = 1.0
    conv = np.fft.irfft(np.fft.rfft(sn) * sb)
    return np.interp(wave, wn, conv, right=np.nan, left=np.nan)


def smooth_R(wave, spec, rsigma, outwave, inres_rsigma):
    """``smoothspec(type='R', outwave=obs, inres=ANN.resolution)``:
    ``smoothing.py:103-115, 132-138, 252-291, 588-608``."""
    sigma_out = CKMS / rsigma
    inres = CKMS / inres_rsigma
    m = _mask(wave, rsigma, outwave)
    w = wave[m]
    s = np.nan_to_num(spec[m], nan=1.0)
    if outwave is None:                                      # smoothing.py:140-141
        outwave = wave
    with np.errstate(invalid='ignore'):
        sigma = np.sqrt(sigma_out ** 2 - inres ** 2)
    wn, sn = _resample_pow2(w, s)
    dv = CKMS * np.median(np.diff(np.log(wn)))
    ss = np.fft.rfftfreq(len(sn), d=dv)
    taper = np.exp(-2 * (np.pi ** 2) * (sigma ** 2) * (ss ** 2))
    conv = np.fft.irfft(np.fft.rfft(sn) * taper)
    return np.interp(outwave, wn, conv, right=np.nan, left=np.nan)


def smooth_lsf(wave, spec, disp, outwave):
    """``smoothspec(type='lsf', fftsmooth=True)``: ``smoothing.py:126-150`` (linear mask of
    ``20 * 100`` AA, ``sigma = resolution[mask]``), ``482-586`` (smooth_lsf_fft) and ``588-608``
    (smooth_fft).  ``disp``: dispersion in AA at every pixel of ``wave``."""
    if outwave is not None:
        wlim = np.array([outwave.min(), outwave.max()])
    else:
        wlim = np.squeeze(np.array([0, np.inf]))
    wlim = wlim + 20.0 * 100 * np.array([-1, 1])
    m = (wave > wlim[0]) & (wave < wlim[1])
    w = wave[m]
    s = np.nan_to_num(spec[m], nan=1.0)
    if outwave is None:
        outwave = wave
    sigma = disp[m]
    dw = np.gradient(w)
    cdf = np.cumsum(dw / sigma)
    cdf /= cdf.max()
    x_per_sigma = np.nanmedian(np.gradient(cdf) / (dw / sigma))
    N = 2 / x_per_sigma
    nx = int(2 ** np.ceil(np.log2(N)))
    x = np.linspace(0, 1, nx)
    dx = 1.0 / nx
    lam = np.interp(x, cdf, w)
    newspec = np.interp(lam, w, s)
    ss = np.fft.rfftfreq(len(newspec), d=dx)
    taper = np.exp(-2 * (np.pi ** 2) * (x_per_sigma ** 2) * (ss ** 2))
    conv = np.fft.irfft(np.fft.rfft(newspec) * taper)
    return np.interp(outwave, lam, conv)


# --------------------------------------------------------------------------- getspec
def getspec(net_fwd, spec, teff, logg, feh, afe, vmic, rot_vel, rad_vel, inst_R,
            outwave, mlp_flux=None, cont=None):
    """``predictspec.py:136-294``.  ``inst_R`` is the sigma-R the caller passes (``genmod.py:82-85``) or,
    when not a float, the LSF vector of ``:265-286``; ``cont = (cont_fwd, cont_spec)`` is the continuum
    emulator of ``:96-102, 208-226``; ``outwave`` may be None (native grid, ``:290-292``)."""
    if np.isfinite(vmic):
        labels = [teff, logg, feh, afe, vmic]              # :188-204
    else:
        labels = [teff, logg, feh, afe]
    modspec = net_fwd(np.asarray(labels)).squeeze() if mlp_flux is None else mlp_flux
    modwave = spec.wavelength
    if cont is not None:                                     # :208-226
        cfwd, cspec = cont
        modcont = cfwd(np.asarray(labels)).squeeze()
        modcont = modcont * (SPEEDOFLIGHT / ((cspec.wavelength * 1E-8) ** 2.0))
        modcont = modcont / np.nanmedian(modcont)
        modspec = modspec * np.interp(modwave, cspec.wavelength, modcont, right=np.nan, left=np.nan)
    if rot_vel != 0.0:                                       # :228-241
        modspec = smooth_vsini(modwave, modspec, rot_vel)
        modspec[0] = modspec[1]
        modspec[-1] = modspec[-2]
    if rad_vel != 0.0:                                       # :243-249
        modwave = modwave * (1.0 + (rad_vel / SPEEDOFLIGHT))
    done = False
    if isinstance(inst_R, float):
        if inst_R > 0.0:                                     # :255-263
            modspec = smooth_R(modwave, modspec, inst_R, outwave, spec.resolution)
            done = True
    else:                                                    # :265-286
        disparr = np.interp(modwave, outwave, inst_R) if outwave is not None else np.asarray(inst_R)
        assert len(disparr) == len(modwave)
        modspec = smooth_lsf(modwave, modspec, disparr, outwave)
        done = True
    if not done and outwave is not None:                     # :288-289
        modspec = np.interp(outwave, modwave, modspec, right=np.nan, left=np.nan)
    return (outwave if outwave is not None else modwave), modspec


def polycalc(coef, inwave):
    """``fitutils.py:11-20``."""
    x = inwave - inwave.min()
    x = 2.0 * (x / x.max()) - 1.0
    return chebval(x, coef)


def genspec(net_fwd, spec, pars, outwave, modpoly, mlp_flux=None):
    """``genmod.py:58-108`` (carbon_bool False)."""
    teff, logg, feh, afe, radvel, rotvel, vmic, inst_R = pars[:8]
    polycoef = pars[8:]
    inst_R = FWHM_FACTOR_FIT * inst_R if isinstance(inst_R, float) else inst_R
    w, f = getspec(net_fwd, spec, teff, logg, feh, afe, vmic, rotvel, radvel,
                   inst_R, outwave, mlp_flux=mlp_flux)
    if modpoly:
        f = f * polycalc(polycoef, w)
    return w, f


# --------------------------------------------------------------------------- photometry
def phot_bc(phot, x):
    """``photANN.py:118-131`` fastANN.encode/eval: fp32 weights against fp64 inputs."""
    xp = ((np.atleast_2d(x) - phot.xmin) / (phot.xmax - phot.xmin)).T
    sig = lambda a: 1.0 / (1 + np.exp(-a))
    a1 = sig(np.matmul(phot.w1, xp) + phot.b1[..., None])
    a2 = sig(np.matmul(phot.w2, a1) + phot.b2[..., None])
    return np.squeeze(np.matmul(phot.w3, a2) + phot.b3[..., None])


def sed(phot, logt, logg, feh, afe, av, rv, logl=None, dist=None, logA=None):
    """``predictsed.py:75-103`` + ``highred.py:19-25``."""
    if av < 5.0:
        BC = phot_bc(phot, [10.0 ** logt, logg, feh, afe, av, rv])
    else:
        BC0 = phot_bc(phot, [10.0 ** logt, logg, feh, afe, 0.0, 3.1])
        a1, b1, a2, b2, c2 = phot.hiav.T
        BC = BC0 - (a1 + b1 * av * (a2 + b2 * rv + c2 * rv ** 2.0))
    if logl is not None and dist is not None:
        return -2.5 * logl + 4.74 - BC + (5.0 * np.log10(dist) - 5.0)
    if logA is not None:
        return 5.0 * logA - 10.0 * (logt - LOG10_TSUN) - 0.26 - BC
    raise IOError('cannot understand input pars into sed function')


def genphot(phot, pars, scaled):
    """``genmod.py:110-155`` (distance mode) / ``:157-187`` (scaled mode); Rv is always
    3.1 because ``likelihood.lnlike`` never passes rvfree (``likelihood.py:103-106``)."""
    teff, logg, feh, afe = pars[:4]
    logt = np.log10(teff)
    if scaled:
        return sed(phot, logt, logg, feh, afe, pars[5], 3.1, logA=pars[4])
    logl = 2.0 * pars[4] + 4.0 * (logt - LOG10_TSUN)
    return sed(phot, logt, logg, feh, afe, pars[6], 3.1, logl=logl, dist=pars[5])


# --------------------------------------------------------------------------- likelihood
SPEC_NAMES = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot', 'Vmic', 'Inst_R']


class OracleLikelihood:
    """``likelihood.py:5-117``: same parameter plumbing, one vector at a time."""

    def __init__(self, cfg, ideal_mlp=False):
        self.cfg = cfg
        self.ideal_mlp = ideal_mlp
        self.spec_bool, self.phot_bool, self.modpoly_bool, self.photscale_bool, _ = cfg.runbools
        self.fitpars_i = list(cfg.fitpars_i)
        self.ndim = len(self.fitpars_i)
        self.fixedpars = dict(cfg.fixedpars)
        self.net = make_net(cfg.spec, ideal=ideal_mlp) if self.spec_bool else None
        self.parsdict = {}

    # likelihood.py:42-82
    def pack(self, pars):
        pd = {pp: vv for pp, vv in zip(self.fitpars_i, pars)}
        pd.update(self.fixedpars)
        self.parsdict = pd
        specpars = photpars = None
        if self.spec_bool:
            specpars = [pd[p] if p in pd else np.nan for p in SPEC_NAMES]
            if self.modpoly_bool:
                specpars = specpars + [pd[p] for p in self.fitpars_i if 'pc' in p]
        if self.phot_bool:
            photpars = [pd[p] for p in ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]']]
            if 'log(A)' in self.fitpars_i:
                photpars += [pd['log(A)']]
            else:
                photpars += [pd['log(R)'], pd['Dist']]
            photpars += [pd['Av']]
            photpars += [pd['Rv'] if 'Rv' in self.fitpars_i else None]
        return specpars, photpars

    def model(self, pars, mlp_flux=None):
        """Model spectrum [n_obs] and magnitudes [nb] for one parameter vector."""
        specpars, photpars = self.pack([float(p) for p in pars])
        flux = mags = None
        if self.spec_bool:
            _, flux = genspec(self.net, self.cfg.spec, specpars, self.cfg.obs_wave,
                              self.modpoly_bool, mlp_flux=mlp_flux)
        if self.phot_bool:
            mags = genphot(self.cfg.phot, photpars, self.photscale_bool)
        return flux, mags

    # likelihood.py:84-117 -- including the per-pixel Python generator of :95-97, which
    # is part of what the reference costs per call
    def lnlikefn(self, pars):
        flux, mags = self.model(pars)
        specchi2 = sedchi2 = 0.0
        if self.spec_bool:
            specchi2 = np.sum([((m - o) ** 2.0) / (s ** 2.0) for m, o, s in
                               zip(flux, self.cfg.obs_flux, self.cfg.obs_eflux)])
        if self.phot_bool:
            op = self.cfg.obs_phot
            sedchi2 = np.sum([((mags[i] - op[kk][0]) ** 2.0) / (op[kk][1] ** 2.0)
                              for i, kk in enumerate(op.keys())])
        return -0.5 * (specchi2 + sedchi2)

    # vectorised bookkeeping for tests: same numbers, no Python pixel loop
    def lnlike_batch(self, theta, return_model=False, batched_mlp=False):
        """Row-by-row by default, i.e. the emulator runs at batch 1 exactly like the
        reference (MKL gemv); ``batched_mlp=True`` runs one fp32 GEMM for all rows, which
        is faster but changes the fp32 summation order (flux differs by ~6e-8 relative)."""
        theta = np.asarray(theta, dtype=np.float64)
        B = theta.shape[0]
        lnl = np.empty(B)
        fluxes = np.empty((B, len(self.cfg.obs_wave))) if self.spec_bool else None
        mags_all = np.empty((B, len(self.cfg.phot.bands))) if self.phot_bool else None
        mlp = None
        if self.spec_bool and batched_mlp:   # ANN.eval is batch-capable (predictspec.py:66-69)
            mlp = self.mlp_batch(theta)
        for i in range(B):
            flux, mags = self.model(theta[i], mlp_flux=None if mlp is None else mlp[i].copy())
            c2 = 0.0
            if self.spec_bool:
                fluxes[i] = flux
                c2 += np.sum(((flux - self.cfg.obs_flux) ** 2.0) / (self.cfg.obs_eflux ** 2.0))
            if self.phot_bool:
                mags_all[i] = mags
                o = np.array([v[0] for v in self.cfg.obs_phot.values()])
                e = np.array([v[1] for v in self.cfg.obs_phot.values()])
                c2 += np.sum(((mags - o) ** 2.0) / (e ** 2.0))
            lnl[i] = -0.5 * c2
        if return_model:
            return lnl, fluxes, mags_all
        return lnl

    def mlp_batch(self, theta):
        """Labels of every row -> fp32 emulator flux [B, D_out] (predictspec.py:188-206)."""
        theta = np.asarray(theta, dtype=np.float64)
        cols = []
        for p in ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]']:
            cols.append(self._col(theta, p))
        vm = self._col(theta, 'Vmic')
        if np.all(np.isfinite(vm)):
            cols.append(vm)
        return self.net(np.stack(cols, axis=1))

    def _col(self, theta, name):
        if name in self.fitpars_i:
            return theta[:, self.fitpars_i.index(name)]
        if name in self.fixedpars:
            return np.full(theta.shape[0], float(self.fixedpars[name]))
        return np.full(theta.shape[0], np.nan)


def model_fn(cfg, theta):
    """Noise-free model at ``theta`` for ``thepayne_b200.synth.build_config``."""
    tmp_flux, tmp_e, tmp_p = cfg.obs_flux, cfg.obs_eflux, cfg.obs_phot
    n = len(cfg.obs_wave)
    cfg.obs_flux, cfg.obs_eflux = np.ones(n), np.ones(n)
    if cfg.phot is not None:
        cfg.obs_phot = {b: [0.0, 1.0] for b in cfg.phot.bands}
    L = OracleLikelihood(cfg)
    _, fl, mg = L.lnlike_batch(theta, return_model=True)
    cfg.obs_flux, cfg.obs_eflux, cfg.obs_phot = tmp_flux, tmp_e, tmp_p
    return fl, mg
