"""The cluster-distributed fast tail (csrc/tail_cluster.cuh: one transform of 32768 / 65536 samples spread over
a thread-block cluster of four CTAs through distributed shared memory) against

  * the golden fixtures minted by the unmodified reference (`mid`: 32768-point transforms, `c4m`: 65536),
  * the single-CTA kernels it replaces (tail_fast.cuh: whole / split transform, one CTA per SM) on the same points,
  * the oracle on the branches the fixtures do not reach with a wide grid: observed pixels in any order
    (np.interp takes them unsorted, smoothing.py:289), a mask so narrow that the second transform fits one CTA
    (the cluster's CTA 0 runs it alone), and no instrumental profile at all (predictspec.py:288-289).

Bars as in test_gpu_parity.py: |dlnL| <= max(1e-3, 1e-8 |lnL|) (|lnL| reaches 2e6 on c4m), flux <= 1e-5 relative.
"""
import copy

import numpy as np
import pytest
import torch

from conftest import load_case
from oracle import payne_oracle as O
from thepayne_b200 import synth

pytestmark = pytest.mark.gpu


def _engine(cfg, prec='parity'):
    from thepayne_b200.engine import engine_from_config
    return engine_from_config(cfg, precision=prec)


def _close(a, b):
    fin = np.isfinite(b)
    return np.array_equal(np.isnan(a), ~fin) and np.all(np.abs(a[fin] - b[fin]) <= np.maximum(1e-3, 1e-8 * np.abs(b[fin])))


@pytest.mark.parametrize('name', ['mid', 'c4m'])
def test_cluster_tail_against_golden_and_single_cta(name):
    cfg, g = load_case(name)
    eng = _engine(cfg)
    # default: on for 65536-sample transforms, off (measured slower than one CTA per SM) for 32768
    assert eng.query('tail_cluster') == (1 if name == 'c4m' else 0) and eng.query('tail_clusters') >= 1
    th = np.concatenate([g['theta'], cfg.draw(40, seed=77)])
    tht = torch.from_numpy(np.ascontiguousarray(th)).cuda()
    res = {}
    for mode in (1, 0):
        eng.set('tail_cluster', mode)
        assert eng.query('tail_cluster') == mode
        flux, _, lnl_m = eng.model_batch(tht)
        lnl = eng.lnlike_batch(tht)
        torch.cuda.synchronize()
        res[mode] = (flux.cpu().numpy(), lnl_m.cpu().numpy(), lnl.cpu().numpy())
        n = len(g['lnl'])
        assert _close(res[mode][1][:n], g['lnl']) and _close(res[mode][2][:n], g['lnl'])      # the reference's values
        nf = g['flux'].shape[0]
        ok = np.isfinite(g['flux'])
        assert np.array_equal(np.isnan(res[mode][0][:nf]), ~ok)
        assert np.max(np.abs(res[mode][0][:nf][ok] - g['flux'][ok]) / np.abs(g['flux'][ok])) < 1e-5
    assert eng.query('status') == 0
    # the two kernels differ only in the order of fp32 butterflies: same spectra to a few ulp, same lnL within the bar
    f1, f0 = res[1][0], res[0][0]
    ok = np.isfinite(f0)
    assert np.array_equal(np.isnan(f1), ~ok)
    assert np.max(np.abs(f1[ok] - f0[ok]) / np.abs(f0[ok])) < 2e-6
    assert _close(res[1][2], res[0][2]) and _close(res[1][1], res[0][1])
    eng.close()


def _wide(model_fn=O.model_fn, **kw):
    args = dict(ann_range=(5100.0, 5400.0), obs_range=(5120.0, 5380.0), n_obs=3000)     # the `mid` shape: N1 = 32768
    args.update(kw)
    return synth.config_mini(model_fn, **args)


def test_cluster_tail_unsorted_pixels():
    """Observed pixels in arbitrary order: every CTA of the cluster then scans all pixels and keeps its own."""
    cfg = _wide()
    perm = np.random.default_rng(5).permutation(len(cfg.obs_wave))
    cfgp = copy.copy(cfg)
    cfgp.obs_wave, cfgp.obs_flux, cfgp.obs_eflux = cfg.obs_wave[perm], cfg.obs_flux[perm], cfg.obs_eflux[perm]
    th = cfg.draw(12, seed=3)
    th[1, cfg.fitpars_i.index('Vrot')] = 0.0
    tht = torch.from_numpy(np.ascontiguousarray(th)).cuda()
    eng, engp = _engine(cfg), _engine(cfgp)
    eng.set('tail_cluster', 1); engp.set('tail_cluster', 1)
    assert engp.query('tail_cluster') == 1
    f, _, l = eng.model_batch(tht)
    fp, _, lp = engp.model_batch(tht)
    l2, lp2 = eng.lnlike_batch(tht), engp.lnlike_batch(tht)
    torch.cuda.synchronize()
    f, fp = f.cpu().numpy(), fp.cpu().numpy()
    np.testing.assert_array_equal(fp, f[:, perm])                       # per-pixel model does not depend on the order
    assert _close(lp.cpu().numpy(), l.cpu().numpy()) and _close(lp2.cpu().numpy(), l2.cpu().numpy())
    ref = O.OracleLikelihood(cfgp).lnlike_batch(th[:4])
    assert _close(lp2.cpu().numpy()[:4], ref)
    eng.close(); engp.close()


def test_cluster_tail_narrow_mask_runs_in_one_cta():
    """N2 <= N1/4: the masked spectrum fits one CTA's buffer; CTA 0 transforms it alone (runtime-planned FFT)."""
    cfg = _wide(obs_range=(5240.0, 5262.0), n_obs=900)
    eng = _engine(cfg)
    eng.set('tail_cluster', 1)
    assert eng.query('tail_cluster') == 1 and eng.query('nfft1') == 32768
    th = cfg.draw(6, seed=9)
    th[0] = cfg.theta_true
    flux, _, lnl = eng.model_batch(torch.from_numpy(np.ascontiguousarray(th)).cuda())
    ref_l, ref_f, _ = O.OracleLikelihood(cfg).lnlike_batch(th, return_model=True)
    assert np.isfinite(ref_f).all()
    assert np.max(np.abs(flux.cpu().numpy() - ref_f) / np.abs(ref_f)) < 1e-5
    assert _close(lnl.cpu().numpy(), ref_l)
    eng.close()


def test_cluster_tail_without_instrumental_profile():
    """Inst_R absent: rotation stage on the cluster, then plain np.interp split over the four CTAs."""
    def mf(cfg, theta):
        keep = [i for i, p in enumerate(cfg.fitpars_i) if p != 'Inst_R']
        if len(keep) != len(cfg.fitpars_i):
            cfg.theta_true = cfg.theta_true[keep]
            cfg.fitpars_i = [cfg.fitpars_i[i] for i in keep]
        return O.model_fn(cfg, cfg.theta_true[None, :])
    cfg = _wide(mf)
    assert 'Inst_R' not in cfg.fitpars_i
    eng = _engine(cfg)
    eng.set('tail_cluster', 1)
    assert eng.query('tail_cluster') == 1
    th = cfg.draw(5, seed=11)
    th[1, cfg.fitpars_i.index('Vrot')] = 0.0
    flux, _, lnl = eng.model_batch(torch.from_numpy(np.ascontiguousarray(th)).cuda())
    ref_l, ref_f, _ = O.OracleLikelihood(cfg).lnlike_batch(th, return_model=True)
    ok = np.isfinite(ref_f)
    f = flux.cpu().numpy()
    assert np.array_equal(np.isnan(f), ~ok)
    assert np.max(np.abs(f[ok] - ref_f[ok]) / np.abs(ref_f[ok])) < 1e-5
    assert _close(lnl.cpu().numpy(), ref_l)
    eng.close()


@pytest.mark.parametrize('name', ['c2', 'mini_joint', 'c4m'])
def test_dynamic_point_scheduling_covers_every_point_once(name):
    """CTAs (clusters) claim their next point from a device counter.  With the persistent grid capped at 3 every
    CTA works through many points of a 41-point batch; each lnL must be the bits of the uncapped run (a point
    processed twice or skipped would show as a stale or missing value), also in the multi-slab case."""
    cfg, g = load_case(name)
    eng = _engine(cfg)
    th = torch.from_numpy(np.ascontiguousarray(np.concatenate([g['theta'][:4], cfg.draw(37, seed=13)]))).cuda()
    ref = eng.lnlike_batch(th).cpu().numpy()
    fref, _, _ = eng.model_batch(th)
    for cap, slab in ((3, 8192), (1, 8192), (2, 16)):
        eng.set('tail_grid_cap', cap)
        eng.set('max_batch', slab)
        got = eng.lnlike_batch(th).cpu().numpy()
        f, _, l = eng.model_batch(th)
        assert np.array_equal(got, ref, equal_nan=True)
        assert np.array_equal(l.cpu().numpy(), ref, equal_nan=True) or _close(l.cpu().numpy(), ref)
        assert np.array_equal(f.cpu().numpy(), fref.cpu().numpy(), equal_nan=True)
    eng.close()
