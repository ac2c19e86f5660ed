"""Numpy model of the in-SMEM FFT convolution the tail kernel runs (dev tool + CPU test).

Real length-N signal -> packed complex length M=N/2 -> in-place DIF passes (natural in,
digit-reversed out) -> pair-wise (k, M-k) filter stage in digit-reversed storage ->
in-place DIT passes (digit-reversed in, natural out).  Mirrors csrc/fft.cuh index math.
"""
import numpy as np


def radix_plan(m):
    """log2 radices, first DIF pass first; last pass is always the contiguous radix-16."""
    assert m >= 8
    rest = m - 4
    plan = []
    while rest > 0:
        if rest % 3 == 0 or rest >= 7 or rest == 3:
            r = 3
        elif rest == 4:
            r = 4
        elif rest in (5,):
            r = 3
        elif rest in (1, 2):
            r = rest
        else:
            r = 3
        # keep L/R >= 16 guaranteed because the last 4 bits are reserved
        plan.append(r)
        rest -= r
    return plan + [4]


def dif_forward(z, plan):
    M = len(z)
    x = z.astype(np.complex128).copy()
    L = M
    for lr in plan:
        R = 1 << lr
        S = L // R
        xb = x.reshape(M // L, R, S)           # [block, m, j]
        q = np.arange(R)
        F = np.exp(-2j * np.pi * np.outer(q, q) / R)      # [q, m]
        y = np.einsum('qm,bmj->bqj', F, xb)
        tw = np.exp(-2j * np.pi * np.outer(q, np.arange(S)) / L)   # W_L^{j q}
        x = (y * tw[None]).reshape(M)
        L = S
    return x


def dit_inverse(x, plan):
    M = len(x)
    x = x.copy()
    Ls = []
    L = M
    for lr in plan:
        Ls.append(L)
        L //= (1 << lr)
    for lr, L in zip(reversed(plan), reversed(Ls)):
        R = 1 << lr
        S = L // R
        xb = x.reshape(M // L, R, S)
        q = np.arange(R)
        tw = np.exp(+2j * np.pi * np.outer(q, np.arange(S)) / L)
        F = np.exp(+2j * np.pi * np.outer(q, q) / R)      # [m, q]
        x = np.einsum('mq,bqj->bmj', F, xb * tw[None]).reshape(M)
    return x / M


def freq_to_pos(k, plan, M):
    """storage position of frequency k after the DIF passes."""
    pos = np.zeros_like(k)
    L = M
    for lr in plan:
        R = 1 << lr
        L //= R
        pos = pos + (k & (R - 1)) * L
        k = k >> lr
    return pos


def filter_pairs(x, H, plan):
    """In-place filter stage on digit-reversed storage; H real, length M+1."""
    M = len(x)
    N = 2 * M
    k = np.arange(M // 2 + 1)
    kp = (M - k) % M
    pk, pkp = freq_to_pos(k, plan, M), freq_to_pos(kp, plan, M)
    Zk, Zp = x[pk], x[pkp]
    E = 0.5 * (Zk + np.conj(Zp))
    O = -0.5j * (Zk - np.conj(Zp))
    W = np.exp(-2j * np.pi * k / N)
    A = 0.5 * (H[k] + H[M - k])
    Bc = 0.5 * (H[k] - H[M - k])
    E2 = A * E + Bc * W * O
    O2 = Bc * np.conj(W) * E + A * O
    out = x.copy()
    out[pk] = E2 + 1j * O2
    sel = kp != k
    out[pkp[sel]] = np.conj(E2[sel]) + 1j * np.conj(O2[sel])
    return out


def conv_real(sig, H):
    N = len(sig)
    M = N // 2
    plan = radix_plan(int(np.log2(M)))
    z = sig[0::2] + 1j * sig[1::2]
    Z = dif_forward(z, plan)
    Z = filter_pairs(Z, H, plan)
    zz = dit_inverse(Z, plan)
    out = np.empty(N)
    out[0::2], out[1::2] = zz.real, zz.imag
    return out


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    for N in [512, 1024, 2048, 4096, 8192, 16384, 32768]:
        M = N // 2
        plan = radix_plan(int(np.log2(M)))
        s = rng.standard_normal(N)
        H = np.exp(-1e-5 * np.arange(M + 1) ** 1.5) * np.cos(np.arange(M + 1) * 0.01)
        ref = np.fft.irfft(np.fft.rfft(s) * H)
        got = conv_real(s, H)
        z = s[0::2] + 1j * s[1::2]
        Zd = dif_forward(z, plan)
        kk = np.arange(M)
        e1 = np.abs(Zd[freq_to_pos(kk, plan, M)] - np.fft.fft(z)).max()
        print(N, plan, 'fft err', e1, 'conv err', np.abs(ref - got).max())


def filter_pairs_odd(x, H, plan):
    """Odd-frequency half of a split transform: local k' <-> frequency 2k'+1 of size M = 2*len(x);
    partner is the complement Mh-1-k'.  Mirrors ct_filter_pairs_odd in csrc/fft_ct.cuh."""
    Mh = len(x)
    M, N = 2 * Mh, 4 * Mh
    kp = np.arange(Mh // 2)
    # enumerate like the kernel: klo in [0, Mlo/2), c in [0,16)
    Mlo = Mh >> 4
    klo, c = np.meshgrid(np.arange(Mlo // 2), np.arange(16), indexing='ij')
    klo, c = klo.ravel(), c.ravel()
    kloc = klo + c * Mlo
    rows = freq_to_pos(klo, plan[:-1], Mlo)
    pk = rows * 16 + c
    pp = (Mlo - 1 - rows) * 16 + (15 - c)
    assert np.array_equal(pp, freq_to_pos(Mh - 1 - kloc, plan, Mh))
    k = 2 * kloc + 1
    Zk, Zp = x[pk], x[pp]
    E = 0.5 * (Zk + np.conj(Zp))
    O = -0.5j * (Zk - np.conj(Zp))
    W = np.exp(-2j * np.pi * k / N)
    A = 0.5 * (H[k] + H[M - k])
    Bc = 0.5 * (H[k] - H[M - k])
    E2 = A * E + Bc * W * O
    O2 = Bc * np.conj(W) * E + A * O
    out = x.copy()
    out[pk] = E2 + 1j * O2
    out[pp] = np.conj(E2) + 1j * np.conj(O2)
    return out


def conv_real_split(sig, H):
    """csrc/fft_ct.cuh::ct_convolve_split: N = 4*Mh real samples, two halves of Mh complex points."""
    N = len(sig)
    M = N // 2
    Mh = M // 2
    plan = radix_plan(int(np.log2(Mh)))
    zz = sig[0::2] + 1j * sig[1::2]
    a, b = zz[:Mh].copy(), zz[Mh:].copy()
    j = np.arange(Mh)
    WM = np.exp(-2j * np.pi * j / M)
    a, b = a + b, (a - b) * WM                       # cross DIF
    Hs = H / M
    a = dit_inverse(filter_pairs(dif_forward(a, plan), Hs[0::2] * Mh, plan), plan) * Mh   # undo model's 1/Mh
    b = dit_inverse(filter_pairs_odd(dif_forward(b, plan), Hs * Mh, plan), plan) * Mh
    a = a / Mh
    b = b / Mh
    b = b * np.conj(WM)                              # cross DIT
    out_z = np.concatenate([a + b, a - b])
    out = np.empty(N)
    out[0::2], out[1::2] = out_z.real, out_z.imag
    return out
