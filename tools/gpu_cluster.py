"""Dev tool: the cluster-distributed tail (tail_cluster.cuh) against the single-CTA split tail on the same points
(golden c4m rows + a drawn batch), with the tail time of both.

Per-phase clocks (profiles/r02_cluster_tail.txt) come from an instrumented build of the one translation unit, linked with
the other objects of a normal build and selected with PAYNE_LIB_PATH:
    cd thepayne_b200/csrc
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -DPAYNE_CLUSTER_PROF \
         -c -o /tmp/tc_prof.o tail_cluster_tu.cu
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o _build/libpayne_prof.so _build/payne_b200.o _build/gemm_tu.o \
         _build/tail_fast_tu.o _build/tail_fast_poly_tu.o _build/tail_general_tu.o /tmp/tc_prof.o
    PAYNE_LIB_PATH=$PWD/_build/libpayne_prof.so python tools/gpu_cluster.py c4m 4096"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_case
from thepayne_b200.engine import engine_from_config

name = sys.argv[1] if len(sys.argv) > 1 else 'c4m'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg, g = load_case(name)
eng = engine_from_config(cfg, precision='parity')
print('cluster', eng.query('tail_cluster'), 'clusters', eng.query('tail_clusters'), 'ctas/sm', eng.query('tail_cluster_ctas_per_sm'),
      'nfft1', eng.query('nfft1'), flush=True)
thg = torch.from_numpy(g['theta']).cuda()
th = torch.from_numpy(cfg.draw(B, seed=5)).cuda()
res = {}
for mode in (1, 0):
    eng.set('tail_cluster', mode)
    flux, mags, lnl = eng.model_batch(thg)
    lg = eng.lnlike_batch(thg)
    torch.cuda.synchronize()
    print('mode', mode, 'golden: max|dlnl| vs ref %.3e (model_batch) %.3e (lnlike_batch); flux rel %.2e' % (
        np.max(np.abs(lnl.cpu().numpy() - g['lnl'])), np.max(np.abs(lg.cpu().numpy() - g['lnl'])),
        np.nanmax(np.abs(flux[:g['flux'].shape[0]].cpu().numpy() - g['flux']) / np.abs(g['flux']))), flush=True)
    eng.set('timing', 1)
    for _ in range(2): out = eng.lnlike_batch(th)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): out = eng.lnlike_batch(th)
    ev1.record(); torch.cuda.synchronize()
    print('mode', mode, 'B=%d: %.3f ms/step, tail %.3f ms, mlp %.3f ms' % (B, ev0.elapsed_time(ev1) / 5, eng.last_ms('tail'), eng.last_ms('mlp')), flush=True)
    eng.set('timing', 0)
    res[mode] = out.cpu().numpy()
    if mode == 1 and hasattr(eng.lib, 'payne_debug_cluster_prof'):
        import ctypes
        buf = (ctypes.c_ulonglong * 32)()
        eng.lib.payne_debug_cluster_prof(buf)          # clear
        out = eng.lnlike_batch(th); torch.cuda.synchronize()
        eng.lib.payne_debug_cluster_prof(buf)
        v = np.array(list(buf), dtype=np.float64)
        names = ['dif+regrid', 'sync', 'fwd', 'sync', 'filter', 'sync', 'inv', 'sync', 'dit', 'sync', 'regrid back', 'sync']
        names = ['s1 ' + x for x in names[:12]] + ['s2 ' + x for x in names[:10]] + ['final', 'reduce+sync']
        tot = v[:24].sum()
        for i, nm in enumerate(names):
            print('  %-16s %6.2f %%   %8.0f clk/point/CTA' % (nm, 100 * v[i] / tot, v[i] / (4 * B)))
        print('  total %.0f clk/point/CTA' % (tot / (4 * B)))
d = np.abs(res[1] - res[0])
print('cluster vs split on %d points: max|dlnl| %.3e, rel %.3e, finite %d/%d' % (B, np.nanmax(d), np.nanmax(d / np.abs(res[0])), np.isfinite(res[1]).sum(), np.isfinite(res[0]).sum()))
