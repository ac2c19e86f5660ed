"""Synthetic datasets for the likelihood hot path.

The reference ships no trained emulators, so every config is rebuilt from seeds:
random-init ``LinNet`` weights of the named architecture (reference
``Payne/train/NNmodels.py:140-152`` -- six ``nn.Linear`` created in the order
lin1..lin6 right after ``torch.manual_seed``), the log-uniform emulator grid of
``Payne/utils/readc3k.py:441-451``, prior-box parameter draws following
``demo/runPayne.py:122-141`` and a mock observation at S/N 50.

Nothing in here touches CUDA; the arrays are plain numpy so that the oracle, the
tests and ``bench.py`` all see bit-identical inputs.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import numpy as np
import torch

SIGMA_TO_FWHM_TRAIN = 2.35482  # Payne/train/trainspec.py:49


@dataclass
class SpecNet:
    """Spectrum emulator container: same fields the reference keeps on ``ANN``
    (``Payne/predict/predictspec.py:43-59``) plus the six weight/bias pairs of
    ``model/lin{1..6}`` (``NNmodels.py:147-152``), fp32, row-major ``[out, in]``."""
    weights: list            # 6 x float32 [out, in]
    biases: list             # 6 x float32 [out]
    xmin: np.ndarray         # float64 [D_in]
    xmax: np.ndarray         # float64 [D_in]
    wavelength: np.ndarray   # float64 [D_out]
    resolution: float        # sigma-R of the emulator grid
    inlabels: list = field(default_factory=lambda: ['teff', 'logg', 'feh', 'afe'])
    encode_offset: float = 0.5   # LinNet subtracts 0.5 (NNmodels.py:166)
    # 'LinNet' (6 x Linear, sigmoid; NNmodels.py:140-162), 'SMLP' (4 x Linear, LeakyReLU, torch fp32;
    # NNmodels.py:92-115), 'YST1' (3 layers, leaky ReLU, numpy fp64; predict/ystpred.py:18-58) or
    # 'MultiNet' (G chunk nets Net(D_in,H,P) of 4 sigmoid/linear layers, train/old/trainspec_multi.py:29-52:
    # weights = [W1 [G,H,D_in], W2 [G,H,H], W3 [G,H,H], W4 [D_out,H]] with the chunks' output layers one
    # after another, biases alike, ``chunk`` = pixels per net, encode_offset 0)
    nntype: str = 'LinNet'
    chunk: int = 0

    @property
    def n_groups(self):
        return int(self.weights[0].shape[0]) if self.nntype == 'MultiNet' else 1

    @property
    def D_in(self):
        return int(self.weights[0].shape[-1])

    @property
    def D_out(self):
        return int(self.weights[-1].shape[0])

    @property
    def n_layers(self):
        return len(self.weights)

    @property
    def activation(self):
        return 'sigmoid' if self.nntype == 'LinNet' else 'leaky'

    def digest(self) -> str:
        h = hashlib.sha256()
        for a in self.weights + self.biases:
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(np.ascontiguousarray(self.wavelength).tobytes())
        return h.hexdigest()[:16]


@dataclass
class PhotNet:
    """Stacked per-band photometry nets, the arrays ``fastANN`` holds
    (``Payne/predict/photANN.py:97-106``)."""
    bands: list
    w1: np.ndarray   # float32 [nb, H, 6]
    b1: np.ndarray   # float32 [nb, H]
    w2: np.ndarray   # float32 [nb, H, H]
    b2: np.ndarray   # float32 [nb, H]
    w3: np.ndarray   # float32 [nb, 1, H]
    b3: np.ndarray   # float32 [nb, 1]
    xmin: np.ndarray  # float64 [6]
    xmax: np.ndarray  # float64 [6]
    hiav: np.ndarray  # float64 [nb, 5]  (a1,b1,a2,b2,c2) of highred.py:10-17, NaN if absent


def ann_wavegrid(w0: float, w1: float, r_fwhm: float):
    """Emulator pixel grid: ``w0*(1+1/(3 R_sigma))**i`` while <= w1
    (``readc3k.py:441-451``), with ``R_sigma = R_fwhm*2.35482``."""
    rsig = r_fwhm * SIGMA_TO_FWHM_TRAIN
    # scalar libm pow in a Python loop, exactly like the trainer: bit-reproducible across hosts
    # (numpy's SIMD pow may differ in the last bit between CPU generations)
    out, i = [], 1
    while True:
        wave_i = w0 * (1.0 + 1.0 / (3.0 * rsig)) ** (i - 1.0)
        if wave_i <= w1:
            out.append(wave_i)
            i += 1
        else:
            break
    w = np.array(out, dtype=np.float64)
    return w, rsig


def make_specnet(D_in, H, wave, resolution, seed=0, out_scale=0.3, out_bias=1.0,
                 H2=None, H3=None, nntype='LinNet'):
    """Random-init LinNet: default torch init, then lin6.weight*=0.3 and
    lin6.bias=1 so the emulated flux looks like a normalised spectrum (1 +/- 0.07).
    ``nntype`` 'SMLP' / 'YST1' build the legacy leaky-ReLU stacks (4 / 3 layers) the same way."""
    H2 = H if H2 is None else H2
    H3 = H if H3 is None else H3
    D_out = len(wave)
    torch.manual_seed(seed)
    dims = {'LinNet': [(D_in, H), (H, H), (H, H2), (H2, H2), (H2, H3), (H3, D_out)],
            'SMLP': [(D_in, H), (H, H2), (H2, H3), (H3, D_out)],          # NNmodels.py:99-107
            'YST1': [(D_in, H), (H, H2), (H2, D_out)]}[nntype]            # ystpred.py:25-30
    lins = [torch.nn.Linear(i, o) for i, o in dims]
    with torch.no_grad():
        lins[-1].weight *= out_scale
        lins[-1].bias.fill_(out_bias)
    xmin = np.array([3500.0, 0.0, -2.5, -0.2, 0.5][:D_in])
    xmax = np.array([8000.0, 5.5, 0.5, 0.6, 3.0][:D_in])
    labels = ['teff', 'logg', 'feh', 'afe', 'vmic'][:D_in]
    return SpecNet(
        weights=[l.weight.detach().numpy().copy() for l in lins],
        biases=[l.bias.detach().numpy().copy() for l in lins],
        xmin=xmin, xmax=xmax, wavelength=np.asarray(wave, dtype=np.float64),
        resolution=float(resolution), inlabels=labels, nntype=nntype)


def make_multinet(D_in, H, wave, resolution, chunk, seed=0, out_scale=0.3, out_bias=1.0):
    """Random-init multi-chunk emulator: one ``Net(D_in, H, P)`` (lin1..lin4 created in that order,
    trainspec_multi.py:29-36) per ``chunk`` pixels, last one narrower; output layers scaled like
    make_specnet's so the flux looks like a normalised spectrum."""
    D_out = len(wave)
    G = (D_out + chunk - 1) // chunk
    torch.manual_seed(seed)
    W1, W2, W3, W4, b1, b2, b3, b4 = [], [], [], [], [], [], [], []
    for g in range(G):
        P = min(chunk, D_out - g * chunk)
        lins = [torch.nn.Linear(D_in, H), torch.nn.Linear(H, H), torch.nn.Linear(H, H), torch.nn.Linear(H, P)]
        with torch.no_grad():
            lins[-1].weight *= out_scale
            lins[-1].bias.fill_(out_bias)
        for lst, l in zip([W1, W2, W3, W4], lins):
            lst.append(l.weight.detach().numpy().copy())
        for lst, l in zip([b1, b2, b3, b4], lins):
            lst.append(l.bias.detach().numpy().copy())
    xmin = np.array([3500.0, 0.0, -2.5, -0.2, 0.5][:D_in])
    xmax = np.array([8000.0, 5.5, 0.5, 0.6, 3.0][:D_in])
    return SpecNet(weights=[np.stack(W1), np.stack(W2), np.stack(W3), np.concatenate(W4, 0)],
                   biases=[np.stack(b1), np.stack(b2), np.stack(b3), np.concatenate(b4, 0)],
                   xmin=xmin, xmax=xmax, wavelength=np.asarray(wave, dtype=np.float64), resolution=float(resolution),
                   inlabels=['teff', 'logg', 'feh', 'afe', 'vmic'][:D_in], encode_offset=0.0, nntype='MultiNet',
                   chunk=int(chunk))


def make_photnet(bands, H=128, seed=7):
    """Random-init per-band ``Net(6,H,1)`` (``photANN.py:21-26``), built band by band
    in list order so the reference harness can recreate the same modules."""
    torch.manual_seed(seed)
    w1, b1, w2, b2, w3, b3 = [], [], [], [], [], []
    for _ in bands:
        l1 = torch.nn.Linear(6, H)
        l2 = torch.nn.Linear(H, H)
        l3 = torch.nn.Linear(H, 1)
        w1.append(l1.weight.detach().numpy().copy()); b1.append(l1.bias.detach().numpy().copy())
        w2.append(l2.weight.detach().numpy().copy()); b2.append(l2.bias.detach().numpy().copy())
        w3.append(l3.weight.detach().numpy().copy()); b3.append(l3.bias.detach().numpy().copy())
    xmin = np.array([2500.0, -1.0, -4.0, -0.2, 0.0, 2.0])
    xmax = np.array([50000.0, 5.5, 0.5, 0.6, 5.0, 5.0])
    from .predict.highred import reference_table
    tab = reference_table()
    hiav = np.array([tab.get(b, (np.nan,) * 5) for b in bands], dtype=np.float64)
    return PhotNet(list(bands), np.array(w1), np.array(b1), np.array(w2), np.array(b2),
                   np.array(w3), np.array(b3), xmin, xmax, hiav)


@dataclass
class SynthConfig:
    """One benchmark / parity configuration: emulators + observation + prior box."""
    name: str
    spec: SpecNet
    obs_wave: np.ndarray
    obs_flux: np.ndarray
    obs_eflux: np.ndarray
    fitpars_i: list          # free-parameter names in sampler order
    fixedpars: dict
    runbools: list           # [spec, phot, modpoly, photscale, carbon]
    box: dict                # name -> (lo, hi)
    theta_true: np.ndarray
    phot: PhotNet | None = None
    obs_phot: dict | None = None   # band -> [mag, err]

    @property
    def ndim(self):
        return len(self.fitpars_i)

    def draw(self, B, seed=1234):
        rng = np.random.default_rng(seed)
        lo = np.array([self.box[p][0] for p in self.fitpars_i])
        hi = np.array([self.box[p][1] for p in self.fitpars_i])
        return lo + (hi - lo) * rng.random((B, self.ndim))


_BOX = {
    'Teff': (4000.0, 8000.0), 'log(g)': (4.0, 5.5), '[Fe/H]': (-0.1, 0.1),
    '[a/Fe]': (-0.1, 0.1), 'Vrad': (-1.0, 1.0), 'Vrot': (0.0, 5.0),
    'Vmic': (0.5, 3.0), 'Inst_R': (30000.0, 37000.0), 'log(A)': (-3.0, 7.0),
    'Av': (0.0, 1.0), 'log(R)': (-0.1, 0.1), 'Dist': (1.0, 200.0),
}
_TRUE = {
    'Teff': 5770.0, 'log(g)': 4.44, '[Fe/H]': 0.0, '[a/Fe]': 0.0, 'Vrad': 0.5,
    'Vrot': 3.0, 'Vmic': 1.0, 'Inst_R': 32000.0, 'log(A)': 1.0, 'Av': 0.5,
    'log(R)': 0.0, 'Dist': 10.0,
}
PROCYON_BANDS = ['Bessell_B', 'Bessell_V', 'Bessell_R', 'Bessell_I',
                 '2MASS_J', '2MASS_H', '2MASS_Ks']


def build_config(name, *, model_fn, ann_range=(5130.0, 5340.0), r_fwhm=50000.0,
                 obs_range=(5150.0, 5320.0), n_obs=7000, H=256, vmic=False,
                 npoly=0, bands=None, photscale=True, photH=128, vrot_max=5.0,
                 snr=50.0, seed_net=0, seed_noise=3, obs_wave=None, nntype='LinNet', chunk=0):
    """Assemble a SynthConfig. ``model_fn(cfg, theta[1,ndim]) -> (flux[1,n_obs], mags)``
    supplies the noiseless model at the truth (the oracle, passed in by the caller so
    that this module has no dependency on ``oracle/``)."""
    wave, rsig = ann_wavegrid(ann_range[0], ann_range[1], r_fwhm)
    if nntype == 'MultiNet':
        spec = make_multinet(5 if vmic else 4, H, wave, rsig, chunk, seed=seed_net)
    else:
        spec = make_specnet(5 if vmic else 4, H, wave, rsig, seed=seed_net, nntype=nntype)
    if obs_wave is None:
        obs_wave = np.linspace(obs_range[0], obs_range[1], n_obs)
    fit = ['Teff', 'log(g)', '[Fe/H]', '[a/Fe]', 'Vrad', 'Vrot']
    if vmic:
        fit.append('Vmic')
    fit.append('Inst_R')
    phot = None
    if bands:
        phot = make_photnet(bands, H=photH)
        fit += (['log(A)'] if photscale else ['log(R)', 'Dist']) + ['Av']
    box = dict(_BOX)
    box['Vrot'] = (0.0, vrot_max)
    true = dict(_TRUE)
    for k in range(npoly):
        fit.append('pc_%d' % k)
        box['pc_%d' % k] = (0.75, 1.25) if k == 0 else (-0.1, 0.1)
        true['pc_%d' % k] = 1.0 if k == 0 else 0.02 * (-1) ** k
    cfg = SynthConfig(
        name=name, spec=spec, obs_wave=np.asarray(obs_wave, dtype=np.float64),
        obs_flux=None, obs_eflux=None, fitpars_i=fit, fixedpars={},
        runbools=[True, bool(bands), npoly > 0, bool(photscale and bands), False],
        box=box, theta_true=np.array([true[p] for p in fit]), phot=phot)
    flux_true, mags_true = model_fn(cfg, cfg.theta_true[None, :])
    rng = np.random.default_rng(seed_noise)
    e = np.abs(flux_true[0]) / snr
    cfg.obs_eflux = e
    cfg.obs_flux = flux_true[0] + e * rng.standard_normal(len(e))
    if bands:
        cfg.obs_phot = {b: [float(m + 0.05 * rng.standard_normal()), 0.05]
                        for b, m in zip(bands, mags_true[0])}
    return cfg


def config_c2(model_fn, **kw):
    """C2: UVES-range solar mock, LinNet 4-256-256-256-14172, n_obs 7000."""
    return build_config('C2-uves', model_fn=model_fn, **kw)


def config_c3(model_fn, **kw):
    """C3: C2 + 7-band photometry, photscale."""
    return build_config('C3-joint', model_fn=model_fn, bands=PROCYON_BANDS, **kw)


def config_c4(model_fn, **kw):
    """C4 (monolithic variant): full-wavelength emulator LinNet 5-512-512-512-51784 (vmic label),
    65536-point transforms, n_obs 25000, order-4 continuum, Vrot up to 100 km/s."""
    args = dict(ann_range=(4750.0, 5500.0), obs_range=(4760.0, 5490.0), n_obs=25000, H=512, vmic=True,
                npoly=5, vrot_max=100.0)
    args.update(kw)
    return build_config('C4-full', model_fn=model_fn, **args)


C1_BANDS = ['2MASS_H', '2MASS_J', '2MASS_Ks', 'Bessell_B', 'Bessell_I', 'Bessell_R', 'Bessell_U', 'Bessell_V',
            'DECam_g', 'DECam_i', 'DECam_r', 'DECam_u', 'DECam_Y', 'DECam_z', 'PS_g', 'PS_i', 'PS_open', 'PS_r',
            'PS_w', 'PS_y', 'PS_z', 'SDSS_g', 'SDSS_i', 'SDSS_r', 'SDSS_u', 'Tycho_B', 'Tycho_V']


def demo_wavegrid(n=25600, w0=5139.250234270761, step=1.0 / 600000.0):
    """Pixel grid of the reference's demo spectrum (demo/demodata.h5 ``spec/wave``: log-uniform,
    5139.25-5363.26 A, 25600 pixels) for runs that do not have the file."""
    return w0 * (1.0 + step) ** np.arange(n)


def config_c1(model_fn, obs_wave=None, obs_flux=None, obs_phot=None, **kw):
    """C1: the reference's mock-Sun demo (demo/runPayne.py:36-143): n_obs 25600 on 5139-5363 A with
    e = flux/25 (runPayne.py:50), 27 photometric bands, emulator LinNet 4-256-256-256-15675 on
    5135-5368 A, parameters (Teff, logg, FeH, aFe, Vrad, Vrot, Inst_R, log(A), Av), photscale.
    ``obs_flux`` / ``obs_phot`` replace the mock observation by the demo file's arrays (the golden
    fixture does this); without them the observation is a mock at the truth like every other config."""
    args = dict(ann_range=(5135.0, 5368.0), bands=C1_BANDS, photH=32, snr=25.0,
                obs_wave=demo_wavegrid() if obs_wave is None else np.asarray(obs_wave, dtype=np.float64))
    args.update(kw)
    cfg = build_config('C1-demo', model_fn=model_fn, **args)
    if obs_flux is not None:
        cfg.obs_flux = np.asarray(obs_flux, dtype=np.float64)
        cfg.obs_eflux = cfg.obs_flux / 25.0
    if obs_phot is not None:
        cfg.obs_phot = {b: [float(obs_phot[b][0]), float(obs_phot[b][1])] for b in cfg.phot.bands}
    return cfg


def config_c4_chunked(model_fn, **kw):
    """C4 (chunked variant): the full-wavelength emulator as G = 13 chunk nets
    Net(5,512,512,512,P=4096) (last chunk 2632 pixels), everything else as config_c4."""
    args = dict(nntype='MultiNet', chunk=4096)
    args.update(kw)
    return config_c4(model_fn, **args)


def config_mini(model_fn, **kw):
    """Small everything: fast on the CPU oracle; used for golden vectors."""
    args = dict(ann_range=(5140.0, 5190.0), obs_range=(5150.0, 5180.0), n_obs=1500, H=64)
    args.update(kw)
    return build_config('mini', model_fn=model_fn, **args)
