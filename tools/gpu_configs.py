"""Run the other BASELINE configs through the engine (dev tool): C3 joint spec+phot at 16k points,
C5-sized slab run (131072 points on one GPU), C4 monolithic (65536-point split transforms)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from oracle import payne_oracle as O
from thepayne_b200.engine import engine_from_config

def run(name, B, ncheck=6):
    cfg, g = load_case(name)
    eng = engine_from_config(cfg, precision='parity')
    th = cfg.draw(B, seed=5)
    tht = torch.from_numpy(th).cuda()
    for _ in range(2): out = eng.lnlike_batch(tht)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): out = eng.lnlike_batch(tht)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    out = out.cpu().numpy()
    L = O.OracleLikelihood(cfg)
    idx = np.linspace(0, B - 1, ncheck).astype(int)
    ref = np.array([L.lnlikefn(th[i]) for i in idx])
    print('%s B=%d: %.2f ms -> %.3e evals/s ; finite %d ; max|dlnL| vs oracle on %d rows %.2e (lnL range %.0f..%.0f) ; mem %.1f GB' % (
        name, B, dt * 1e3, B / dt, int(np.isfinite(out).sum()), ncheck, np.max(np.abs(out[idx] - ref)), ref.min(), ref.max(),
        torch.cuda.max_memory_allocated() / 1e9), flush=True)
    eng.close()

run('c3', 16384)
run('c2', 131072)
run('c4m', 4096, ncheck=3)
