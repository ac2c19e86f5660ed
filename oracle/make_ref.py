"""Stage the UNMODIFIED reference hot-path files under oracle/_ref/ (build container only).

TEST / BENCH INFRASTRUCTURE.  The reference is pure Python; ``/root/reference`` does not exist on
the GPU box, so ``bench.py --impl reference`` could only time the oracle port there.  This recipe
copies the files the likelihood path executes byte for byte into ``oracle/_ref/Payne``
-- a git-ignored directory that travels with the gpurun snapshot like a built .so -- so that the
reference arm and the ``cpu_baseline`` leg run the reference itself (``oracle/refharness.py`` points at
whichever tree exists).  Nothing is edited: MANIFEST.json records the sha256 of every source and copy.

    python -m oracle.make_ref            # also run by __graft_entry__.build() when /root/reference exists
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('PAYNE_REFERENCE', '/root/reference')

# what lnlikefn -> genspec -> getspec -> smoothspec -> chi2 (+ sed) imports, plus the prior the
# sampler adapter is pinned against and the legacy / multi-chunk networks the goldens are minted with
FILES = [
    'Payne/utils/smoothing.py',
    'Payne/utils/quantiles.py',
    'Payne/train/NNmodels.py',
    'Payne/train/old/trainspec_multi.py',
    'Payne/predict/predictspec.py',
    'Payne/predict/photANN.py',
    'Payne/predict/highred.py',
    'Payne/predict/predictsed.py',
    'Payne/predict/ystpred.py',
    'Payne/fitting/fitutils.py',
    'Payne/fitting/genmod.py',
    'Payne/fitting/likelihood.py',
    'Payne/fitting/prior.py',
    'Payne/fitting/advancedpriors.py',
]


def sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def main():
    if not os.path.isdir(os.path.join(SRC, 'Payne')):
        print('make_ref: no reference tree at %s; leaving %s as it is' % (SRC, DST))
        return 0 if os.path.isdir(DST) else 1
    man = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        assert sha(s) == sha(d)
        man[rel] = sha(d)
    json.dump({'source': SRC, 'sha256': man}, open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1)
    print('make_ref: %d reference files staged under %s' % (len(FILES), DST))
    return 0


if __name__ == '__main__':
    sys.exit(main())
