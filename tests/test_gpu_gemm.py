"""GPU: the tcgen05 GEMM kernel alone (payne_gemm_test hook) against float64 matmul.

The parity mode ("X3": three 8-bit fixed-point bf16 slices per operand) must be EXACT whenever
the dominant slice products sum to an fp32-representable integer -- that is the property the
whole scheme rests on (the TMEM accumulator never rounds) -- and correctly rounded to fp32
accuracy on random data."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gemm(A, W, b, prec):
    from thepayne_b200 import _lib
    lib = _lib.load()
    A, W, b = [np.ascontiguousarray(x, dtype=np.float32) for x in (A, W, b)]
    M, K = A.shape
    N = W.shape[0]
    C = np.empty((M, N), dtype=np.float32)
    _lib.check(lib.payne_gemm_test(A.ctypes.data, W.ctypes.data, b.ctypes.data, M, N, K, _lib.PREC[prec], 0,
                                   C.ctypes.data))
    return C


@pytest.mark.parametrize('K', [64, 256, 512])
def test_x3_integer_sums_are_exact(K):
    rng = np.random.default_rng(K)
    M, N = 300, 700                                   # ragged in M and N on purpose
    A = rng.integers(0, 256, (M, K)).astype(np.float64) / 256.0          # one slice: p1 only
    W = rng.integers(-128, 129, (N, K)).astype(np.float64) / 128.0 * 0.125
    W[:, 0] = 0.125 * (1 - 2.0 ** -7)                 # pins every row scale to 2^-3
    b = np.zeros(N)
    C = gemm(A, W, b, 'parity')
    ref = A @ W.T                                     # exact in float64 (integers < 2^53)
    assert np.array_equal(C.astype(np.float64), ref)


@pytest.mark.parametrize('prec,bar', [('parity', 1.5e-7), ('3xtf32', 3e-6), ('tf32', 3e-3)])
def test_gemm_accuracy_random(prec, bar):
    rng = np.random.default_rng(7)
    M, N, K = 513, 1000, 256
    A = 1.0 / (1.0 + np.exp(-rng.standard_normal((M, K))))               # sigmoid-like activations
    W = (rng.random((N, K)) - 0.5) * 0.04
    b = rng.standard_normal(N) * 0.1 + 1.0
    A32, W32, b32 = A.astype(np.float32), W.astype(np.float32), b.astype(np.float32)
    ref = A32.astype(np.float64) @ W32.astype(np.float64).T + b32.astype(np.float64)
    C = gemm(A32, W32, b32, prec)
    err = np.abs(C - ref) / np.abs(ref)
    assert err.max() <= bar, err.max()
    if prec == 'parity':
        # no coherent bias: the mean signed error is far below one fp32 ulp
        assert abs(np.mean((C - ref) / np.abs(ref))) < 2e-9
