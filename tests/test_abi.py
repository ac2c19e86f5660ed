"""CPU: the C-ABI library builds, loads, and exports every symbol include/payne_b200.h declares.
No compute calls are made (there is no GPU here); creating a context must fail LOUDLY, not fall
back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from thepayne_b200 import _lib, build
    build.build()            # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    return _lib.load()


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'payne_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(payne_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported(lib):
    from thepayne_b200 import _lib
    names = declared_functions()
    assert set(names) == set(_lib.EXPORTS), (names, _lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'payne_b200.h')).read()
    assert lib.payne_abi_version() == int(re.search(r'#define\s+PAYNE_ABI_VERSION\s+(\d+)', hdr).group(1))


def test_struct_layout_matches_header(lib, tmp_path):
    """ctypes mirrors of the structs agree with what a C compiler makes of the header."""
    import subprocess
    from thepayne_b200 import _lib
    src = tmp_path / 's.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "payne_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(PayneSpecNet), sizeof(PaynePhotNet),'
                   'sizeof(PayneObs), sizeof(PayneLayout), offsetof(PayneLayout, fixed), offsetof(PayneLayout, precision));'
                   'return 0;}\n')
    exe = tmp_path / 's'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.PayneSpecNet), ctypes.sizeof(_lib.PaynePhotNet), ctypes.sizeof(_lib.PayneObs),
            ctypes.sizeof(_lib.PayneLayout), _lib.PayneLayout.fixed.offset, _lib.PayneLayout.precision.offset]
    assert got == want
    assert _lib.NPAR == 13 and _lib.PAR_INDEX['Rv'] == 12


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from thepayne_b200 import _lib
    lay = _lib.PayneLayout()
    lay.spec_bool, lay.ndim = 1, 7
    ob = _lib.PayneObs()
    sp = _lib.PayneSpecNet()
    ctx = ctypes.c_void_p()
    rc = lib.payne_ctx_create(ctypes.byref(sp), None, ctypes.byref(ob), ctypes.byref(lay), 0, ctypes.byref(ctx))
    assert rc == -2 and not ctx.value
    assert b'no CPU fallback' in lib.payne_last_error()
    from thepayne_b200.engine import Engine
    with pytest.raises(_lib.PayneError):
        Engine(spec=None, fitpars_i=['Teff'])


def test_product_never_imports_oracle():
    """The package must not route through oracle/ (tier rule): no import of it anywhere."""
    pkg = os.path.join(ROOT, 'thepayne_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith('.py'):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, os.path.join(dp, f)


def test_build_units_and_their_headers_exist():
    """thepayne_b200/build.py: every translation unit and every header its staleness check lists is in csrc/ (a
    header renamed without updating the table would silently stop triggering rebuilds)."""
    from thepayne_b200 import build as b
    for unit, hdrs in b.UNITS.items():
        assert os.path.exists(os.path.join(b.CSRC, unit)), unit
        for h in hdrs:
            assert os.path.exists(os.path.join(b.CSRC, h)), (unit, h)
    import glob
    on_disk = {os.path.basename(p) for p in glob.glob(os.path.join(b.CSRC, '*.cu'))}
    assert on_disk == set(b.UNITS), (on_disk, set(b.UNITS))


def test_peer_gather_needs_a_process_group():
    from thepayne_b200 import dist as pdist
    with pytest.raises(RuntimeError):
        pdist.PeerGather(None, 16)
    a = pdist._DevArray(0x1000, 8).__cuda_array_interface__
    assert a['shape'] == (8,) and a['typestr'] == '<f8' and a['data'] == (0x1000, False)
