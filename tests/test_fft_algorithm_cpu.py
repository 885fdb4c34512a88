"""The arithmetic of the hand-written FFT passes (incompact3d_b200/csrc/x3d_fft_kernels.cuh), restated in numpy and checked
against numpy.fft on the CPU: the Stockham stage (which points a thread holds, where they go, which twiddles they take), the
per-stage twiddle tables of the contiguous pass (x3d_fft.cu: build_stage_tables), the half-length trick of the real z transforms
with its untangling step, and the unnormalised conventions of 2DECOMP&FFT (forward sign -1, inverse +1; the caller divides,
src/poisson.f90:333).  The kernels themselves are compared with the oracle on the GPU (tests/test_poisson_gpu.py)."""
import numpy as np
import pytest

RADICES = {16: (8, 2), 32: (8, 4), 64: (8, 8), 128: (8, 8, 2), 256: (8, 8, 4), 512: (8, 8, 8), 1024: (8, 8, 8, 2)}   # FftRadix<N>


def dft_small(v, inverse):
    """radix-R butterfly = R-point DFT of the R inputs"""
    r = len(v)
    s = +1 if inverse else -1
    w = np.exp(s * 2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
    return w @ v


def stage_tables(n):
    """TW_s[t-1][b] = exp(-2 pi i t (b mod NS) / (NS R)) for the stages after the first (build_stage_tables)"""
    tabs = []
    ns = RADICES[n][0]
    for r in RADICES[n][1:]:
        b = np.arange(n // r)
        tabs.append(np.array([np.exp(-2j * np.pi * t * (b % ns) / (ns * r)) for t in range(1, r)]))
        ns *= r
    return tabs


def stockham(x, inverse=False, use_stage_tables=False):
    """fft_run: before every stage thread j holds the points j + m T (m = 0..7, T = N/8); for radix R these are 8/R butterflies
    b = j + u T with inputs at slots u + t (8/R); outputs go to rows (b / NS) NS R + (b mod NS) + t NS"""
    n = len(x)
    T = n // 8
    W = np.exp(-2j * np.pi * np.arange(n) / n)   # the single table of the strided passes
    tabs = stage_tables(n)
    data = np.array(x, dtype=complex)
    ns = 1
    for si, r in enumerate(RADICES[n]):
        q = 8 // r
        out = np.empty(n, dtype=complex)
        for j in range(T):
            v = [data[j + m * T] for m in range(8)]          # the thread's registers
            for u in range(q):
                b = j + u * T
                k = b % ns
                pts = []
                for t in range(r):
                    val = v[u + t * q]
                    if ns > 1 and t > 0:
                        w = tabs[si - 1][t - 1][b] if use_stage_tables else W[t * k * (n // (ns * r))]
                        val = val * (np.conj(w) if inverse else w)
                    pts.append(val)
                res = dft_small(np.array(pts), inverse)
                j0 = (b // ns) * (ns * r) + k
                for t in range(r):
                    out[j0 + t * ns] = res[t]
        data = out
        ns *= r
    return data


@pytest.mark.parametrize("n", sorted(RADICES))
@pytest.mark.parametrize("tables", [False, True])
def test_stockham_stages_give_the_dft(n, tables):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.allclose(stockham(x, False, tables), np.fft.fft(x), rtol=0, atol=1e-11 * n)
    assert np.allclose(stockham(x, True, tables), np.fft.ifft(x) * n, rtol=0, atol=1e-11 * n)   # unnormalised inverse


@pytest.mark.parametrize("n", sorted(RADICES))
def test_last_stage_leaves_the_read_pattern(n):
    """after the last stage slot m of thread j holds point j + m T -- what a first stage reads: the fused x pass goes from the
    forward to the inverse transform without re-ordering"""
    T = n // 8
    r = RADICES[n][-1]
    ns = n // r
    q = 8 // r
    for j in range(T):
        for u in range(q):
            b = j + u * T
            j0 = (b // ns) * (ns * r) + b % ns
            for t in range(r):
                assert j0 + t * ns == j + (u + t * q) * T


@pytest.mark.parametrize("n", [32, 64, 128, 512])
def test_real_transforms_by_the_half_length_trick(n):
    """k_fft_z_r2c / k_fft_z_c2r: M = n/2 complex points z[m] = x[2m] + i x[2m+1]"""
    rng = np.random.default_rng(3 * n)
    x = rng.standard_normal(n)
    M = n // 2
    U = np.exp(-2j * np.pi * np.arange(M) / n)
    Z = stockham(x[0::2] + 1j * x[1::2])
    X = np.empty(M + 1, dtype=complex)
    for k in range(M):
        a, b = Z[k], np.conj(Z[(M - k) % M])
        E, O = 0.5 * (a + b), (a - b) / 2j
        X[k] = E + U[k] * O
        if k == 0:
            X[M] = (E - O).real
    assert np.allclose(X, np.fft.rfft(x), rtol=0, atol=1e-11 * n)
    # inverse, unnormalised: Z[k] = (X[k] + conj X[M-k]) + i (X[k] - conj X[M-k]) conj(U[k])
    Zi = np.array([(X[k] + np.conj(X[M - k])) + 1j * (X[k] - np.conj(X[M - k])) * np.conj(U[k]) for k in range(M)])
    z = stockham(Zi, inverse=True)
    back = np.empty(n)
    back[0::2], back[1::2] = z.real, z.imag
    assert np.allclose(back, n * x, rtol=0, atol=1e-10 * n)   # = numpy.fft.irfft(X) * n
