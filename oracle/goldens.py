"""Golden-vector case list shared by ``oracle/make_golden.py`` (which runs the unmodified
reference in the build container) and ``tests/`` (which replay the cases on the oracle and
on the CUDA path).  TEST INFRASTRUCTURE.

Each case = synthetic-config kwargs (networks are re-created from seeds, their sha256
digest is stored in the fixture) + a deterministic parameter batch that plants the edge
cases the reference branches on:
  * Vrot == 0  -> rotational stage skipped            (predictspec.py:229-231)
  * Vrad == 0  -> no Doppler shift                    (predictspec.py:244-245)
  * Inst_R absent/NaN -> plain np.interp fallback      (likelihood.py:51-55, predictspec.py:288)
  * requested resolution above the emulator's -> NaN   (smoothing.py:271)
  * observed pixels outside emulator coverage -> NaN   (smoothing.py:289)
  * Av >= 5 -> analytic high-extinction branch         (predictsed.py:86-90)
"""
import numpy as np

from thepayne_b200 import synth

CASES = {
    # name: (builder kwargs, n_points, n_flux_rows_stored)
    'mini_spec': (dict(kind='mini'), 24, 24),
    'mini_noinst': (dict(kind='mini', drop=['Inst_R'], fixed={'Vrot': 2.5}), 8, 8),
    'mini_joint': (dict(kind='mini', npoly=3, bands=synth.PROCYON_BANDS, photH=16), 16, 16),
    'mini_dist': (dict(kind='mini', bands=synth.PROCYON_BANDS[:3], photH=16, photscale=False,
                       vmic=True), 8, 8),
    'mini_edge': (dict(kind='mini', obs_range=(5139.0, 5160.0), n_obs=600), 4, 4),
    'c2': (dict(kind='c2'), 48, 6),
    'c3': (dict(kind='c3'), 16, 2),
    # 32768-point transforms (one CTA per SM) and 65536-point split transforms (C4-shaped)
    'mid': (dict(kind='mini', ann_range=(5100.0, 5400.0), obs_range=(5120.0, 5380.0), n_obs=3000), 12, 3),
    'c4m': (dict(kind='c4'), 8, 2),
    # awkward shapes: hidden width 100 (not a multiple of 8), odd pixel counts
    'mini_odd': (dict(kind='mini', H=100, ann_range=(5141.0, 5187.3), obs_range=(5150.0, 5180.0), n_obs=1501), 8, 8),
    # legacy leaky-ReLU emulators: SMLP = torch fp32, 4 layers; YST1 = numpy fp64, 3 layers
    'mini_smlp': (dict(kind='mini', nntype='SMLP'), 12, 12),
    'mini_yst': (dict(kind='mini', nntype='YST1', H=48), 12, 12),
    # multi-chunk emulators (Payne/train/old/trainspec_multi.py): chunk widths that are / are not multiples
    # of 32 pixels (one grouped launch / one launch per chunk for the output layers), and C4's chunked variant
    'mini_multi': (dict(kind='mini', nntype='MultiNet', chunk=1024, H=64), 12, 12),
    'mini_multi_odd': (dict(kind='mini', nntype='MultiNet', chunk=500, H=40, vmic=True), 8, 8),
    'c4c': (dict(kind='c4c'), 6, 1),
    # C1: the reference's own demo (demo/runPayne.py:36-143): the spectrum and magnitudes of
    # demo/demodata.h5 (n_obs 25600, e = flux/25, 27 bands of which two have no high-Av coefficients)
    # against a random-init emulator of the demo's shape; the fixture carries the observation
    'c1': (dict(kind='c1'), 8, 1),
}
DEMO_H5 = 'demo/demodata.h5'


def demo_observation(ref_root):
    """(obs_wave, obs_flux, {band: [mag, err]}) of the reference's demo file, read without h5py."""
    import os
    from thepayne_b200 import h5lite
    d = h5lite.read(os.path.join(ref_root, DEMO_H5))
    names = [x.decode('ascii') if isinstance(x, bytes) else str(x) for x in d['phot/filter']]
    mags = dict(zip(names, d['phot/phot']))
    return d['spec/wave'], d['spec/flux'], {b: [float(mags[b]), 0.05] for b in synth.C1_BANDS}


def build(name, model_fn, data=None):
    """``data``: for 'c1' the demo observation -- a loaded fixture (tests) or ``demo_observation()``."""
    kw = dict(CASES[name][0])
    kind = kw.pop('kind')
    if kind == 'c1':
        if data is None:
            raise ValueError("case 'c1' needs the demo observation (fixture or demo_observation())")
        if isinstance(data, tuple):
            wave, flux, phot = data
        else:
            wave, flux = data['obs_wave'], data['obs_flux']
            phot = {b: [float(v[0]), float(v[1])] for b, v in zip(synth.C1_BANDS, data['obs_phot'])}
        return synth.config_c1(model_fn, obs_wave=wave, obs_flux=flux, obs_phot=phot, **kw)
    drop = kw.pop('drop', [])
    fixed = kw.pop('fixed', {})
    base = {'mini': synth.config_mini, 'c2': synth.config_c2, 'c3': synth.config_c3, 'c4': synth.config_c4,
            'c4c': synth.config_c4_chunked}[kind]
    if not drop and not fixed:
        return base(model_fn, **kw)

    # build with a wrapper so that the truth used for the mock observation already has the
    # dropped / fixed parameters applied
    def mf(cfg, theta):
        _apply(cfg, drop, fixed)
        return model_fn(cfg, cfg.theta_true[None, :])
    cfg = base(mf, **kw)
    return cfg


def _apply(cfg, drop, fixed):
    keep = [i for i, p in enumerate(cfg.fitpars_i) if p not in drop and p not in fixed]
    if len(keep) != len(cfg.fitpars_i):
        cfg.theta_true = cfg.theta_true[keep]
        cfg.fitpars_i = [cfg.fitpars_i[i] for i in keep]
        cfg.fixedpars = dict(fixed)


def thetas(name, cfg):
    n = CASES[name][1]
    th = cfg.draw(n, seed=4321)
    th[0] = cfg.theta_true
    ix = {p: i for i, p in enumerate(cfg.fitpars_i)}

    def put(row, par, val):
        if par in ix and row < n:
            th[row, ix[par]] = val
    put(1, 'Vrot', 0.0)
    put(2, 'Vrad', 0.0)
    put(3, 'Vrot', 0.0); put(3, 'Vrad', 0.0)
    put(4, 'Inst_R', 60000.0)            # finer than the emulator -> NaN
    put(5, 'Av', 6.0)                    # high-extinction branch
    put(6, 'Vrot', 35.0)
    put(7, 'Vrad', -40.0)
    if name == 'mini_spec':
        put(8, 'Inst_R', 12000.0)        # broad kernel
        put(9, 'Inst_R', 48900.0)        # sub-pixel kernel (sigma ~ 0.6 px)
        put(10, 'Vrot', 0.3)
        put(11, 'Vrad', 250.0)           # mask endpoints move by ~40 px
    return th


# --------------------------------------------------------------------------- getspec-only branches
def getspec_case():
    """Inputs of the ``f3_getspec`` fixture: the two ``getspec`` branches the likelihood never reaches --
    the continuum emulator (``Cnnpath``, predictspec.py:96-102, 208-226) and the LSF vector
    (predictspec.py:265-286).  Returns (spec, cont, calls); every call is a keyword dictionary for
    ``PayneSpecPredict.getspec`` plus ``use_cont``."""
    wave, rsig = synth.ann_wavegrid(5140.0, 5190.0, 50000.0)
    spec = synth.make_specnet(4, 64, wave, rsig, seed=0)
    # the continuum net has its own, coarser grid that stops short of the red end of the spectrum's:
    # pixels beyond it are NaN after the multiply (np.interp right=nan)
    cwave = np.linspace(5135.0, 5178.0, 301)
    cont = synth.make_specnet(4, 32, cwave, rsig, seed=5)
    outwave = np.linspace(5150.0, 5180.0, 1500)
    lsf = 0.070 + 0.012 * (outwave - 5150.0) / 30.0 + 0.002 * np.sin((outwave - 5150.0) / 3.0)
    wave_r = wave * (1.0 + (12.0 / 299792.458))
    lsf_native = 0.080 + 0.010 * (wave_r - wave_r[0]) / (wave_r[-1] - wave_r[0])
    sun = {'Teff': 5770.0, 'log(g)': 4.44, '[Fe/H]': 0.0, '[a/Fe]': 0.0}
    cool = {'Teff': 4600.0, 'log(g)': 4.7, '[Fe/H]': -0.08, '[a/Fe]': 0.05}
    calls = [
        dict(sun, rot_vel=3.0, rad_vel=0.5, inst_R=32000.0 * 2.355, outwave=outwave, use_cont=True),
        dict(cool, rot_vel=0.0, rad_vel=-25.0, outwave=outwave, use_cont=True),            # plain interp: NaN red end
        dict(cool, rot_vel=8.0, rad_vel=0.0, inst_R=30000.0 * 2.355, outwave=None, use_cont=True),
        dict(sun, rot_vel=3.0, rad_vel=0.5, inst_R=lsf, outwave=outwave, use_cont=False),
        dict(cool, rot_vel=0.0, rad_vel=0.0, inst_R=lsf, outwave=outwave, use_cont=False),
        dict(sun, rot_vel=15.0, rad_vel=-60.0, inst_R=1.4 * lsf, outwave=outwave, use_cont=True),
        dict(sun, rot_vel=2.0, rad_vel=12.0, inst_R=lsf_native, outwave=None, use_cont=False),
    ]
    return spec, cont, calls
