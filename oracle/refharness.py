"""Run the UNMODIFIED reference likelihood path from /root/reference (build container only).

TEST INFRASTRUCTURE.  The reference cannot be imported normally here: h5py, astropy,
dynesty and matplotlib are absent and ``Payne/__init__.py`` pulls all of them in.  The
recipe: register empty stand-ins for the missing third-party modules,
register *shell* packages whose ``__path__`` points into the reference tree (so no
``__init__.py`` runs), import the hot-path modules as they are, and inject random-init
networks into instances made with ``__new__`` (their constructors only read HDF5).

``/root/reference`` does not exist on the GPU box; there the same unmodified files are found
under ``oracle/_ref`` (staged byte for byte by ``oracle/make_ref.py`` in the build container,
git-ignored, shipped with the gpurun snapshot).  Consumers: ``oracle/make_golden.py`` (writes
``tests/golden/*.npz``), ``bench.py --impl reference`` and bench.py's ``cpu_baseline`` leg.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
REF_ROOT = os.environ.get('PAYNE_REFERENCE') or (
    '/root/reference' if os.path.isdir('/root/reference/Payne') else _STAGED)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'Payne'))


def _ascii_read(text, **kw):
    """10-line stand-in for astropy.io.ascii.read: whitespace table -> structured array
    (only ``highred.py:169`` uses it)."""
    rows = [ln.split() for ln in text.strip().splitlines() if ln.strip()]
    names, rows = rows[0], rows[1:]
    dt = [(names[0], 'U32')] + [(n, 'f8') for n in names[1:]]
    return np.array([tuple([r[0]] + [float(v) for v in r[1:]]) for r in rows], dtype=dt)


_mods = None


def load():
    """Import the reference hot-path modules; returns a namespace of them."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    for name in ['h5py', 'dynesty', 'astropy', 'astropy.io']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    # advancedpriors.py:14-25 imports a few astropy names at module level; none is used by the
    # pv_* transforms or the additive gaussian/uniform priors exercised here
    for name in ['astropy.utils', 'astropy.utils.exceptions', 'astropy.units', 'astropy.coordinates']:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['astropy.utils.exceptions'].AstropyWarning = type('AstropyWarning', (Warning,), {})
    sys.modules['astropy.utils.exceptions'].AstropyDeprecationWarning = type('AstropyDeprecationWarning', (Warning,), {})
    sys.modules['astropy'].units = sys.modules['astropy.units']
    sys.modules['astropy.coordinates'].SkyCoord = object
    sys.modules['astropy.coordinates'].CylindricalRepresentation = object
    asc = types.ModuleType('astropy.io.ascii')
    asc.read = _ascii_read
    sys.modules['astropy.io.ascii'] = asc
    sys.modules['astropy.io'].ascii = asc
    sys.modules['astropy'].io = sys.modules['astropy.io']
    pk = types.ModuleType('Payne')
    pk.__path__ = [os.path.join(REF_ROOT, 'Payne')]
    pk.__abspath__ = os.path.join(REF_ROOT, 'Payne') + '/'
    sys.modules['Payne'] = pk
    for sub in ['fitting', 'predict', 'train', 'utils']:
        m = types.ModuleType('Payne.' + sub)
        m.__path__ = [os.path.join(REF_ROOT, 'Payne', sub)]
        sys.modules['Payne.' + sub] = m
        setattr(pk, sub, m)
    # train/old/trainspec_multi.py (the multi-chunk Net) sits one level deeper than when it was written:
    # its `from ..utils.pullspectra import pullspectra` now resolves to Payne.train.utils, which does not
    # exist; the trainer-only helper is never called on the prediction path
    old = types.ModuleType('Payne.train.old')
    old.__path__ = [os.path.join(REF_ROOT, 'Payne', 'train', 'old')]
    sys.modules['Payne.train.old'] = old
    tu = types.ModuleType('Payne.train.utils')
    tu.__path__ = []
    sys.modules['Payne.train.utils'] = tu
    ps = types.ModuleType('Payne.train.utils.pullspectra')
    ps.pullspectra = object
    sys.modules['Payne.train.utils.pullspectra'] = ps
    ns = types.SimpleNamespace()
    for short, full in [('smoothing', 'Payne.utils.smoothing'), ('NNmodels', 'Payne.train.NNmodels'),
                        ('predictspec', 'Payne.predict.predictspec'), ('photANN', 'Payne.predict.photANN'),
                        ('highred', 'Payne.predict.highred'), ('predictsed', 'Payne.predict.predictsed'),
                        ('ystpred', 'Payne.predict.ystpred'), ('fitutils', 'Payne.fitting.fitutils'), ('genmod', 'Payne.fitting.genmod'),
                        ('likelihood', 'Payne.fitting.likelihood'), ('prior', 'Payne.fitting.prior'),
                        ('trainspec_multi', 'Payne.train.old.trainspec_multi')]:
        setattr(ns, short, importlib.import_module(full))
    _mods = ns
    return ns


def build_likelihood(cfg):
    """Reference ``likelihood`` object wired to the synthetic emulators of ``cfg``."""
    R = load()
    s = cfg.spec
    like = R.likelihood.likelihood.__new__(R.likelihood.likelihood)
    like.verbose = False
    like.fitargs = {'obs_wave_fit': cfg.obs_wave, 'obs_flux_fit': cfg.obs_flux,
                    'obs_eflux_fit': cfg.obs_eflux, 'fixedpars': dict(cfg.fixedpars)}
    (like.spec_bool, like.phot_bool, like.modpoly_bool, like.photscale_bool,
     like.carbon_bool) = cfg.runbools
    like.fixedpars = like.fitargs['fixedpars']
    like.fitpars_i = list(cfg.fitpars_i)
    like.ndim = len(like.fitpars_i)
    GM = R.genmod.GenMod()
    like.GM = GM
    if like.spec_bool and getattr(s, 'nntype', 'LinNet') == 'YST1':
        # ystpred.Net holds plain arrays (ystpred.py:25-37); PayneSpecPredict as GenMod builds it for
        # NNtype='YST1' (genmod.py:18-21)
        net = R.ystpred.Net.__new__(R.ystpred.Net)
        net.w_array_0, net.w_array_1, net.w_array_2 = [w.copy() for w in s.weights]
        net.b_array_0, net.b_array_1, net.b_array_2 = [b.copy() for b in s.biases]
        net.xmin, net.xmax = s.xmin.copy(), s.xmax.copy()
        net.wavelength = s.wavelength.copy()
        net.resolution = float(s.resolution)
        PP = R.ystpred.PayneSpecPredict.__new__(R.ystpred.PayneSpecPredict)
        PP.anns, PP.Canns, PP.NN, PP.NNtype = net, None, {}, 'YST1'
        GM.PP = PP
    elif like.spec_bool and getattr(s, 'nntype', 'LinNet') == 'MultiNet':
        # The reference tree has no predictor for its own multi-chunk trainer output; the natural one is an
        # ANN whose ``model`` runs every chunk's reference ``Net`` (trainspec_multi.py:29-67, built the way
        # its readNN does, :717-737) and joins the outputs.  Everything downstream (ANN.eval, getspec,
        # smoothspec, likelihood) is the unmodified reference.
        class _Chunks(object):
            def __init__(self, nets, D_in):
                self.nets, self.D_in = nets, D_in

            def __call__(self, xvar):
                return torch.cat([n(xvar) for n in self.nets], dim=-1)
        nets = []
        H = s.weights[0].shape[1]
        for g in range(s.n_groups):
            lo, hi = g * s.chunk, min((g + 1) * s.chunk, s.D_out)
            net = R.trainspec_multi.Net(s.D_in, H, hi - lo)
            net.xmin, net.xmax = s.xmin, s.xmax
            sd = {}
            for k in range(3):
                sd['lin%d.weight' % (k + 1)] = torch.from_numpy(s.weights[k][g].copy())
                sd['lin%d.bias' % (k + 1)] = torch.from_numpy(s.biases[k][g].copy())
            sd['lin4.weight'] = torch.from_numpy(s.weights[3][lo:hi].copy())
            sd['lin4.bias'] = torch.from_numpy(s.biases[3][lo:hi].copy())
            net.load_state_dict(sd)
            net.eval()
            nets.append(net)
        ann = R.predictspec.ANN.__new__(R.predictspec.ANN)
        ann.model, ann.wavelength = _Chunks(nets, s.D_in), s.wavelength.copy()
        ann.resolution = np.array(s.resolution, dtype=float)
        ann.xmin, ann.xmax, ann.inlabels, ann.NNtype = s.xmin, s.xmax, s.inlabels, 'MultiNet'
        PP = R.predictspec.PayneSpecPredict.__new__(R.predictspec.PayneSpecPredict)
        PP.anns, PP.Canns, PP.NN, PP.NNtype = ann, None, {}, 'MultiNet'
        GM.PP = PP
    elif like.spec_bool:
        nntype = getattr(s, 'nntype', 'LinNet')
        if nntype == 'SMLP':
            H1, H2, H3 = [w.shape[0] for w in s.weights[:3]]
            model = R.NNmodels.SMLP(s.D_in, H1, H2, H3, s.D_out, s.xmin, s.xmax)
            sd = {}
            for k in range(4):
                sd['features.%d.weight' % (2 * k)] = torch.from_numpy(s.weights[k].copy())
                sd['features.%d.bias' % (2 * k)] = torch.from_numpy(s.biases[k].copy())
        else:
            H1, H2, H3 = s.weights[0].shape[0], s.weights[3].shape[0], s.weights[4].shape[0]
            model = R.NNmodels.LinNet(s.D_in, H1, H2, H3, s.D_out, s.xmin, s.xmax)
            sd = {}
            for k in range(6):
                sd['lin%d.weight' % (k + 1)] = torch.from_numpy(s.weights[k].copy())
                sd['lin%d.bias' % (k + 1)] = torch.from_numpy(s.biases[k].copy())
        model.load_state_dict(sd)
        model.eval()
        model.D_in = s.D_in
        ann = R.predictspec.ANN.__new__(R.predictspec.ANN)
        ann.model, ann.wavelength = model, s.wavelength.copy()
        ann.resolution = np.array(s.resolution, dtype=float)
        ann.xmin, ann.xmax, ann.inlabels, ann.NNtype = s.xmin, s.xmax, s.inlabels, nntype
        PP = R.predictspec.PayneSpecPredict.__new__(R.predictspec.PayneSpecPredict)
        PP.anns, PP.Canns, PP.NN, PP.NNtype = ann, None, {}, nntype
        GM.PP = PP
    if like.phot_bool:
        p = cfg.phot
        like.fitargs['obs_phot'] = cfg.obs_phot
        H = p.w1.shape[1]
        nnlist = []
        for b in range(len(p.bands)):
            net = R.photANN.Net(6, H, 1)
            net.load_state_dict({
                'lin1.weight': torch.from_numpy(p.w1[b].copy()), 'lin1.bias': torch.from_numpy(p.b1[b].copy()),
                'lin2.weight': torch.from_numpy(p.w2[b].copy()), 'lin2.bias': torch.from_numpy(p.b2[b].copy()),
                'lin3.weight': torch.from_numpy(p.w3[b].copy()), 'lin3.bias': torch.from_numpy(p.b3[b].copy())})
            net.xmin, net.xmax = p.xmin, p.xmax
            nnlist.append(types.SimpleNamespace(model=net))
        fpp = R.predictsed.FastPayneSEDPredict.__new__(R.predictsed.FastPayneSEDPredict)
        fpp.filternames = list(p.bands)
        fpp.anns = R.photANN.fastANN(nnlist, fpp.filternames)
        fpp.HiAv = R.highred.highAv(fpp.filternames)
        GM.fppsed, GM.filterarray = fpp, list(p.bands)
    return like


def _ref_ann(R, s):
    """Reference ``ANN`` (predictspec.py:29-74) around a random-init LinNet / SMLP container."""
    nntype = getattr(s, 'nntype', 'LinNet')
    if nntype == 'SMLP':
        H1, H2, H3 = [w.shape[0] for w in s.weights[:3]]
        model = R.NNmodels.SMLP(s.D_in, H1, H2, H3, s.D_out, s.xmin, s.xmax)
        sd = {}
        for k in range(4):
            sd['features.%d.weight' % (2 * k)] = torch.from_numpy(s.weights[k].copy())
            sd['features.%d.bias' % (2 * k)] = torch.from_numpy(s.biases[k].copy())
    else:
        H1, H2, H3 = s.weights[0].shape[0], s.weights[3].shape[0], s.weights[4].shape[0]
        model = R.NNmodels.LinNet(s.D_in, H1, H2, H3, s.D_out, s.xmin, s.xmax)
        sd = {}
        for k in range(6):
            sd['lin%d.weight' % (k + 1)] = torch.from_numpy(s.weights[k].copy())
            sd['lin%d.bias' % (k + 1)] = torch.from_numpy(s.biases[k].copy())
    model.load_state_dict(sd)
    model.eval()
    model.D_in = s.D_in
    ann = R.predictspec.ANN.__new__(R.predictspec.ANN)
    ann.model, ann.wavelength = model, s.wavelength.copy()
    ann.resolution = np.array(s.resolution, dtype=float)
    ann.xmin, ann.xmax, ann.inlabels, ann.NNtype = s.xmin, s.xmax, s.inlabels, nntype
    return ann


def ref_getspec(spec, cont, calls):
    """Reference ``PayneSpecPredict.getspec`` (predictspec.py:136-294) with an optional continuum emulator
    (``Canns``, :96-102); ``calls`` = list of keyword dictionaries, returns the list of (wave, flux)."""
    R = load()
    PP = R.predictspec.PayneSpecPredict.__new__(R.predictspec.PayneSpecPredict)
    PP.anns, PP.NN, PP.NNtype = _ref_ann(R, spec), {}, getattr(spec, 'nntype', 'LinNet')
    PP.Canns = _ref_ann(R, cont) if cont is not None else None
    out = []
    for kw in calls:
        w, f = PP.getspec(**kw)
        out.append((np.array(w, dtype=np.float64), np.array(f, dtype=np.float64)))
    return out


def ref_model(cfg, theta):
    """Reference model spectrum/mags for each row of theta (via genspec/genphot*)."""
    like = build_likelihood(cfg)
    fl, mg = [], []
    for t in np.asarray(theta, dtype=np.float64):
        like.lnlikefn_pack = None
        # replay lnlikefn's packing (likelihood.py:42-72) by calling it with a dummy lnlike
        captured = {}
        orig = like.lnlike
        like.lnlike = lambda specpars=None, photpars=None: captured.update(s=specpars, p=photpars) or 0.0
        like.lnlikefn(t)
        like.lnlike = orig
        if like.spec_bool:
            _, f = like.GM.genspec(captured['s'], outwave=cfg.obs_wave, modpoly=like.modpoly_bool)
            fl.append(np.array(f, dtype=np.float64))
        if like.phot_bool:
            d = (like.GM.genphot_scaled if like.photscale_bool else like.GM.genphot)(captured['p'])
            mg.append(np.array([d[b] for b in cfg.phot.bands], dtype=np.float64))
    return (np.array(fl) if fl else None), (np.array(mg) if mg else None)


def ref_model_fn(cfg, theta):
    """``model_fn`` for synth.build_config driven by the reference itself."""
    tmp = (cfg.obs_flux, cfg.obs_eflux, cfg.obs_phot)
    n = len(cfg.obs_wave)
    cfg.obs_flux, cfg.obs_eflux = np.ones(n), np.ones(n)
    if cfg.phot is not None:
        cfg.obs_phot = {b: [0.0, 1.0] for b in cfg.phot.bands}
    out = ref_model(cfg, theta)
    cfg.obs_flux, cfg.obs_eflux, cfg.obs_phot = tmp
    return out


def ref_prior(inpriordict, names, free, runbools, U, fixed=None):
    """Reference prior.priortrans / lnpriorfn for each row of the unit-cube sample U."""
    R = load()
    flags = {n: n in free for n in names}
    P = R.prior.prior({'fixedpars': dict(fixed or {})}, inpriordict, [names, flags], runbools)
    theta = np.array([P.priortrans(list(u)) for u in U], dtype=np.float64)
    lnp = np.array([P.lnpriorfn({k: v for k, v in zip(P.fitpars_i, t)}) for t in theta])
    return theta, lnp


def ref_lnlike(cfg, theta):
    like = build_likelihood(cfg)
    return np.array([like.lnlikefn(t) for t in np.asarray(theta, dtype=np.float64)])


def vectorise_chi2(like):
    """Replace the reference's Python-generator chi2 (likelihood.py:95-97, ~40 % of its time) by the
    numpy expression a maintainer would write -- the 'vectorised-chi2 variant' of SURVEY.md §8d.  The
    model path is untouched; only the two sums of likelihood.py:94-112 change."""
    def lnlike(specpars=None, photpars=None):
        specchi2 = sedchi2 = 0.0
        if like.spec_bool:
            _, m = like.GM.genspec(specpars, outwave=like.fitargs['obs_wave_fit'], modpoly=like.modpoly_bool,
                                   carbon_bool=like.carbon_bool)
            r = (np.asarray(m) - like.fitargs['obs_flux_fit']) / like.fitargs['obs_eflux_fit']
            specchi2 = np.sum(r * r)
        if like.phot_bool:
            sed = like.GM.genphot_scaled(photpars) if like.photscale_bool else like.GM.genphot(photpars)
            sedchi2 = np.sum([((sed[k] - v[0]) ** 2.0) / (v[1] ** 2.0) for k, v in like.fitargs['obs_phot'].items()])
        return -0.5 * (specchi2 + sedchi2)
    like.lnlike = lnlike
    return like
