"""Leaky-ReLU emulators (SMLP / YST1) at C2 size: emulator time with the output layer on the tensor cores
(precision 'parity') against all CUDA-core layers ('simt'), and the two against each other -- dev tool."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thepayne_b200 import synth
from thepayne_b200.engine import engine_from_config


def model_fn(cfg, theta):
    n = len(cfg.obs_wave)
    cfg.obs_flux, cfg.obs_eflux = np.ones(n), np.ones(n)
    eng = engine_from_config(cfg, precision='simt')
    fl, mg, _ = eng.model_batch(torch.from_numpy(np.ascontiguousarray(theta)).cuda())
    eng.close()
    return fl.cpu().numpy(), None


B = 4096
for nntype in ['YST1', 'SMLP']:
    cfg = synth.config_c2(model_fn, nntype=nntype, H=256)
    th = torch.from_numpy(np.ascontiguousarray(cfg.draw(B, seed=1))).cuda()
    res = {}
    for prec in ['parity', 'simt']:
        eng = engine_from_config(cfg, precision=prec)
        eng.set('timing', 1)
        for _ in range(3): out = eng.lnlike_batch(th)
        t = []
        for _ in range(10):
            out = eng.lnlike_batch(th); torch.cuda.synchronize(); t.append((eng.last_ms('mlp'), eng.last_ms('tail')))
        res[prec] = out.cpu().numpy()
        print('%s %-6s D_out %d: emulator %.3f ms, tail %.3f ms' % (nntype, prec, cfg.spec.D_out, np.median([a for a, _ in t]),
                                                                   np.median([b for _, b in t])), flush=True)
        eng.close()
    d = np.abs(res['parity'] - res['simt'])
    print('   |lnL(parity) - lnL(simt)| max %.2e at |lnL| up to %.1e' % (np.nanmax(d), np.nanmax(np.abs(res['simt']))))
