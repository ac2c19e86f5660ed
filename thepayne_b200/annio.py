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


def save_specnet(path, net: SpecNet):
    d = {'label_i': np.array([s.encode() for s in net.inlabels]), 'xmin': net.xmin, 'xmax': net.xmax,
         'wavelengths': net.wavelength, 'resolution': np.array(net.resolution)}
    for k in range(6):
        d['model/lin%d.weight' % (k + 1)] = net.weights[k]
        d['model/lin%d.bias' % (k + 1)] = net.biases[k]
    if path.endswith('.h5'):
        from . import h5lite
        return h5lite.write(path, d, gzip=('lin',))      # trainflux.py:567-570 gzips every model tensor
    np.savez_compressed(path, **d)


def load_specnet(path) -> SpecNet:
    d = _open(path)
    labels = [x.decode('utf-8') if isinstance(x, bytes) else str(x) for x in d['label_i']]
    return SpecNet(weights=[np.asarray(d['model/lin%d.weight' % k], dtype=np.float32) for k in range(1, 7)],
                   biases=[np.asarray(d['model/lin%d.bias' % k], dtype=np.float32) for k in range(1, 7)],
                   xmin=np.asarray(d['xmin'], dtype=np.float64), xmax=np.asarray(d['xmax'], dtype=np.float64),
                   wavelength=np.asarray(d['wavelengths'], dtype=np.float64),
                   resolution=float(np.asarray(d['resolution'], dtype=float)), inlabels=labels)


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
