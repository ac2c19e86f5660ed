import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


_cache = {}


def load_case(name):
    """Rebuild a golden case: networks from seeds (digest-checked), observation, thetas and
    the reference outputs from tests/golden/<name>.npz."""
    if name in _cache:
        return _cache[name]
    from oracle import goldens, payne_oracle
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    cfg = goldens.build(name, payne_oracle.model_fn, data=g if name == 'c1' else None)
    assert cfg.spec.digest() == str(g['digest']), 'seeded network differs from the fixture'
    assert list(g['fitpars']) == cfg.fitpars_i
    cfg.obs_flux, cfg.obs_eflux = g['obs_flux'], g['obs_eflux']
    np.testing.assert_array_equal(cfg.obs_wave, g['obs_wave'])
    if 'obs_phot' in g:
        cfg.obs_phot = {b: [float(v[0]), float(v[1])] for b, v in zip(cfg.phot.bands, g['obs_phot'])}
    _cache[name] = (cfg, g)
    return cfg, g
