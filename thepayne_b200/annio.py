"""Emulator containers on disk.

The reference stores a trained spectrum ANN as HDF5 with datasets ``label_i, xmin, xmax,
wavelengths, resolution, model/lin{1..6}.{weight,bias}`` (Payne/train/trainspec.py:214-228,
read by Payne/predict/predictspec.py:43-59 and Payne/train/NNmodels.py:44-89) and one
``nnMIST_{band}.h5`` per photometric band with ``model/lin{1,2,3}.*, xmin, xmax``
(Payne/predict/photANN.py:60-80).  Those files are read with h5py when it is installed and
with the pure-Python reader ``h5lite`` otherwise; the same dataset names are also accepted from
an ``.npz`` archive (``convert_h5``).  ``save_specnet`` / ``save_photnet`` write ``.h5`` in the
reference's layout when the path ends in ``.h5``.
"""
from __future__ import annotations

import os

import numpy as np

from .synth import PhotNet, SpecNet


def _open(path):
    if path.endswith('.npz'):
        z = np.load(path, allow_pickle=False)
        return {k: z[k] for k in z.files}
    try:
        import h5py
    except ImportError:
        # no libhdf5 in this image: the reference's files (h5py defaults) are within what the
        # pure-Python reader understands; anything else raises a descriptive IOError
        from . import h5lite
        return h5lite.read(path)
    out = {}
    with h5py.File(path, 'r') as f:
        def visit(name, obj):
            if hasattr(obj, 'shape'):
                out[name] = obj[()]
        f.visititems(visit)
    return out


def convert_h5(src, dst):
    d = _open(src)
    np.savez_compressed(dst, **d)
    return dst


def _spec_datasets(net: SpecNet):
    """Dataset names per network type, as the reference's trainers / readers use them."""
    if net.nntype == 'YST1':                       # predict/ystpred.py:25-37
        d = {'x_min': net.xmin, 'x_max': net.xmax, 'wavelength': net.wavelength,
             'resolution': np.array([net.resolution])}
        for k in range(3):
            d['w_array_%d' % k] = net.weights[k]
            d['b_array_%d' % k] = net.biases[k]
        return d
    d = {'label_i': np.array([s.encode() for s in net.inlabels]), 'xmin': net.xmin, 'xmax': net.xmax,
         'wavelengths': net.wavelength, 'resolution': np.array(net.resolution)}
    for k in range(net.n_layers):
        name = 'model/lin%d' % (k + 1) if net.nntype == 'LinNet' else 'model/features.%d' % (2 * k)  # NNmodels.py:51-63
        d[name + '.weight'] = net.weights[k]
        d[name + '.bias'] = net.biases[k]
    return d


def save_specnet(path, net: SpecNet):
    d = _spec_datasets(net)
    if path.endswith('.h5'):
        from . import h5lite
        return h5lite.write(path, d, gzip=('lin', 'features'))      # trainflux.py:567-570 gzips every model tensor
    np.savez_compressed(path, **d)


def load_specnet(path, NNtype=None) -> SpecNet:
    """``NNtype`` None: inferred from the dataset names (LinNet ``model/lin*``, SMLP
    ``model/features.*``, YST1 ``w_array_*``)."""
    d = _open(path)
    if NNtype is None:
        NNtype = 'YST1' if 'w_array_0' in d else ('SMLP' if 'model/features.0.weight' in d else 'LinNet')
    f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    if NNtype == 'YST1':
        return SpecNet(weights=[f32(d['w_array_%d' % k]) for k in range(3)],
                       biases=[f32(d['b_array_%d' % k]) for k in range(3)],
                       xmin=f64(d['x_min']), xmax=f64(d['x_max']), wavelength=f64(d['wavelength']),
                       resolution=float(np.asarray(d['resolution'], dtype=float).reshape(-1)[0]), nntype='YST1')
    labels = [x.decode('utf-8') if isinstance(x, bytes) else str(x) for x in d['label_i']]
    names = ['model/lin%d' % k for k in range(1, 7)] if NNtype == 'LinNet' else \
        ['model/features.%d' % k for k in (0, 2, 4, 6)]
    return SpecNet(weights=[f32(d[n + '.weight']) for n in names], biases=[f32(d[n + '.bias']) for n in names],
                   xmin=f64(d['xmin']), xmax=f64(d['xmax']), wavelength=f64(d['wavelengths']),
                   resolution=float(np.asarray(d['resolution'], dtype=float)), inlabels=labels, nntype=NNtype)


def save_photnet(dirpath, net: PhotNet, fmt='npz'):
    os.makedirs(dirpath, exist_ok=True)
    for i, b in enumerate(net.bands):
        if fmt == 'h5':
            from . import h5lite
            h5lite.write(os.path.join(dirpath, 'nnMIST_%s.h5' % b), {
                'model/lin1.weight': net.w1[i], 'model/lin1.bias': net.b1[i],
                'model/lin2.weight': net.w2[i], 'model/lin2.bias': net.b2[i],
                'model/lin3.weight': net.w3[i], 'model/lin3.bias': net.b3[i],
                'xmin': net.xmin, 'xmax': net.xmax}, gzip=('lin',))
            continue
        np.savez_compressed(os.path.join(dirpath, 'nnMIST_%s.npz' % b), **{
            'model/lin1.weight': net.w1[i], 'model/lin1.bias': net.b1[i],
            'model/lin2.weight': net.w2[i], 'model/lin2.bias': net.b2[i],
            'model/lin3.weight': net.w3[i], 'model/lin3.bias': net.b3[i],
            'xmin': net.xmin, 'xmax': net.xmax})


def load_photnet(dirpath, bands, hiav=None) -> PhotNet:
    acc = {k: [] for k in ['w1', 'b1', 'w2', 'b2', 'w3', 'b3']}
    xmin = xmax = None
    for b in bands:
        p = os.path.join(dirpath, 'nnMIST_%s.npz' % b)
        if not os.path.exists(p):
            p = os.path.join(dirpath, 'nnMIST_%s.h5' % b)
        d = _open(p)
        for k, n in [('w1', 'lin1.weight'), ('b1', 'lin1.bias'), ('w2', 'lin2.weight'),
                     ('b2', 'lin2.bias'), ('w3', 'lin3.weight'), ('b3', 'lin3.bias')]:
            acc[k].append(np.asarray(d['model/' + n], dtype=np.float32))
        if xmin is None:   # fastANN takes the limits of the first band (photANN.py:106-111)
            xmin, xmax = np.asarray(d['xmin'], dtype=np.float64), np.asarray(d['xmax'], dtype=np.float64)
    if hiav is None:
        from .predict.highred import highAv
        hiav = np.array(highAv(bands).Avlist, dtype=np.float64)
    return PhotNet(list(bands), *[np.array(acc[k]) for k in ['w1', 'b1', 'w2', 'b2', 'w3', 'b3']],
                   xmin, xmax, np.asarray(hiav, dtype=np.float64))


# ---------------------------------------------------------------------------------------------
# Multi-chunk emulator files (Payne/train/old/trainspec_multi.py).  The trainer writes one file per
# chunk net, ``{prefix}_w{wavestart}_{waveend}.h5`` (:300-303), holding ``wavelength`` (the chunk's
# pixels, :304-305) and the state dict under ``model_{wavestart}_{waveend}/model/lin{1..4}.{weight,bias}``
# (gzip, :308-313).  Its reader takes the label limits from the caller (readNN(nnpath, wavestart,
# wavestop, xmin, xmax), :717-737); so does this one, together with the emulator's sigma-resolution.
def _chunk_path(prefix, w0, w1):
    return '{0}_w{1}_{2}.h5'.format(prefix, w0, w1)


def save_multinet(prefix, net: SpecNet):
    """Write ``net`` (nntype 'MultiNet') as the trainer would: one .h5 per chunk; returns the paths."""
    from . import h5lite
    assert net.nntype == 'MultiNet'
    paths = []
    for g in range(net.n_groups):
        lo, hi = g * net.chunk, min((g + 1) * net.chunk, net.D_out)
        w = net.wavelength[lo:hi]
        grp = 'model_{0}_{1}/model/'.format(w[0], w[-1])
        d = {'wavelength': w}
        for k in range(3):
            d[grp + 'lin%d.weight' % (k + 1)] = net.weights[k][g]
            d[grp + 'lin%d.bias' % (k + 1)] = net.biases[k][g]
        d[grp + 'lin4.weight'] = net.weights[3][lo:hi]
        d[grp + 'lin4.bias'] = net.biases[3][lo:hi]
        path = _chunk_path(prefix, w[0], w[-1])
        h5lite.write(path, d, gzip=('lin',))
        paths.append(path)
    return paths


def load_multinet(paths, xmin, xmax, resolution, inlabels=None) -> SpecNet:
    """Chunk files (any order; a glob pattern or a list of paths) -> one 'MultiNet' SpecNet with the
    chunks sorted by wavelength.  Every chunk but the last must have the same number of pixels."""
    import glob
    if isinstance(paths, str):
        paths = sorted(glob.glob(paths))
    if not paths:
        raise IOError('no multi-chunk emulator files found')
    chunks = []
    for p in paths:
        d = _open(p)
        wave = np.asarray(d['wavelength'], dtype=np.float64)
        pre = [k[:-len('lin1.weight')] for k in d if k.endswith('model/lin1.weight')]
        if len(pre) != 1:
            raise IOError('%s: expected exactly one model_*/model group' % p)
        g = pre[0]
        W = [np.ascontiguousarray(np.asarray(d[g + 'lin%d.weight' % k], dtype=np.float32)) for k in range(1, 5)]
        b = [np.ascontiguousarray(np.asarray(d[g + 'lin%d.bias' % k], dtype=np.float32)) for k in range(1, 5)]
        if W[3].shape[0] != len(wave):
            raise IOError('%s: lin4 has %d outputs for %d pixels' % (p, W[3].shape[0], len(wave)))
        chunks.append((wave[0], wave, W, b))
    chunks.sort(key=lambda c: c[0])
    P = len(chunks[0][1])
    if any(len(c[1]) != P for c in chunks[:-1]) or len(chunks[-1][1]) > P:
        raise IOError('chunks must hold the same number of pixels (the last one may be narrower)')
    xmin, xmax = np.asarray(xmin, dtype=np.float64), np.asarray(xmax, dtype=np.float64)
    D_in = chunks[0][2][0].shape[1]
    return SpecNet(weights=[np.stack([c[2][k] for c in chunks]) for k in range(3)] + [np.concatenate([c[2][3] for c in chunks], 0)],
                   biases=[np.stack([c[3][k] for c in chunks]) for k in range(3)] + [np.concatenate([c[3][3] for c in chunks], 0)],
                   xmin=xmin, xmax=xmax, wavelength=np.concatenate([c[1] for c in chunks]), resolution=float(resolution),
                   inlabels=list(inlabels) if inlabels else ['teff', 'logg', 'feh', 'afe', 'vmic'][:D_in],
                   encode_offset=0.0, nntype='MultiNet', chunk=int(P))
