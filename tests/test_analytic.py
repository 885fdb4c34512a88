"""Analytic known answers for the compact operators: pins that depend on neither the oracle's code nor the
transpiled Fortran, only on the published schemes (Lele 1992; coefficient values in src/schemes.f90:443-1066).

 * modified wavenumbers: a periodic compact operator maps a Fourier mode to the same mode times a closed-form factor
   (first derivative 6th order, second derivative 6th order, staggered 6th-order derivative in both directions);
 * the same for the symmetric / antisymmetric boundary variants (_11 with npaire 1 / 0) on cosine / sine modes, which
   are the periodic operators on the mirrored line;
 * polynomial exactness of the one-sided closures (_22): every row of the system is at least third-order, so cubics
   are differentiated exactly.

The same checks run on the oracle (CPU) and on the CUDA library through the C ABI (GPU)."""
import numpy as np
import pytest

import helpers as H
import oracle_lib as ol

N, LEN = 64, 2.0 * np.pi
PERIODIC_MODES = [1, 2, 5, 13, 24, 31]


def _apply(backend, x3d, name, u, A, npaire, axis):
    if backend == "oracle":
        return H.oracle_op(name, u, A, npaire)
    H.configure(x3d, A, axis)
    return H.product_op(x3d, name, u, A, npaire)


def _field(line, axis, shape=(3, 2)):
    """a 3-D array whose lines along `axis` all equal `line` (times a lane-dependent factor)"""
    n = len(line)
    dims = [shape[0], shape[1]]
    dims.insert(axis, n)
    lanes = 1.0 + 0.25 * np.arange(shape[0] * shape[1]).reshape(shape)
    sl = [None, None]
    sl.insert(axis, slice(None))
    lane_idx = [slice(None), slice(None)]
    lane_idx.insert(axis, None)
    return np.asfortranarray(line[tuple(sl)] * lanes[tuple(lane_idx)]), lanes[tuple(lane_idx)]


def _kprime_d1(k, h):     # alpha = 1/3, a = 14/9, b = 1/9
    return ((14.0 / 9.0) * np.sin(k * h) + (1.0 / 18.0) * np.sin(2 * k * h)) / (h * (1.0 + (2.0 / 3.0) * np.cos(k * h)))


def _kprime2_d2(k, h):    # alpha = 2/11, a = 12/11, b = 3/11
    return ((24.0 / 11.0) * (1 - np.cos(k * h)) + (3.0 / 22.0) * (1 - np.cos(2 * k * h))) / (h * h * (1.0 + (4.0 / 11.0) * np.cos(k * h)))


def _kprime_stag(k, h):   # alpha = 9/62, a = 63/62, b = 17/62
    return ((63.0 / 31.0) * np.sin(0.5 * k * h) + (17.0 / 93.0) * np.sin(1.5 * k * h)) / (h * (1.0 + (9.0 / 31.0) * np.cos(k * h)))


def check_periodic_modified_wavenumbers(backend, x3d=None, tol=2e-13):
    h = LEN / N
    worst = 0.0
    for axis, ax in enumerate("xyz"):
        A = ol.Axis(N, 0, 0, LEN)
        xs = np.arange(N) * h
        xm = xs + 0.5 * h
        for m in PERIODIC_MODES:
            k = 2 * np.pi * m / LEN
            u, lanes = _field(np.sin(k * xs), axis)
            cosv, _ = _field(np.cos(k * xs), axis)
            d1 = _apply(backend, x3d, f"der{ax}_00", u, A, 1, axis)
            worst = max(worst, np.abs(d1 - _kprime_d1(k, h) * cosv).max() / (k * lanes.max()))
            d2 = _apply(backend, x3d, f"der{ax}{ax}_00", u, A, 1, axis)
            worst = max(worst, np.abs(d2 + _kprime2_d2(k, h) * u).max() / (k * k * lanes.max()))
            cosm, _ = _field(np.cos(k * xm), axis)
            dvp = _apply(backend, x3d, f"der{ax}vp", u, A, 0, axis)       # value at x_i -> derivative at x_{i+1/2}
            worst = max(worst, np.abs(dvp - _kprime_stag(k, h) * cosm).max() / (k * lanes.max()))
            sinm, _ = _field(np.sin(k * xm), axis)
            cosv2, _ = _field(np.cos(k * xs), axis)
            dpv = _apply(backend, x3d, f"der{ax}pv", sinm, A, 1, axis)    # value at x_{i+1/2} -> derivative at x_i
            worst = max(worst, np.abs(dpv - _kprime_stag(k, h) * cosv2).max() / (k * lanes.max()))
    assert worst < tol, worst


def check_symmetric_modes(backend, x3d=None, tol=2e-13):
    """_11 variants: npaire = 1 differentiates an even function (cos), npaire = 0 an odd one (sin); nodes include both
    ends, h = L / (n - 1); the result equals the periodic operator on the mirrored line of 2 (n - 1) points"""
    n = 33
    h = LEN / (n - 1)
    xs = np.arange(n) * h
    worst = 0.0
    for axis, ax in enumerate("xyz"):
        A = ol.Axis(n, 1, 1, LEN)
        for m in (1, 3, 7, 12):
            k = np.pi * m / LEN
            even, lanes = _field(np.cos(k * xs), axis)
            odd, _ = _field(np.sin(k * xs), axis)
            d = _apply(backend, x3d, f"der{ax}_11", even, A, 1, axis)
            worst = max(worst, np.abs(d + _kprime_d1(k, h) * odd).max() / (k * lanes.max()))
            d = _apply(backend, x3d, f"der{ax}_11", odd, A, 0, axis)
            worst = max(worst, np.abs(d - _kprime_d1(k, h) * even).max() / (k * lanes.max()))
            d = _apply(backend, x3d, f"der{ax}{ax}_11", even, A, 1, axis)
            worst = max(worst, np.abs(d + _kprime2_d2(k, h) * even).max() / (k * k * lanes.max()))
            d = _apply(backend, x3d, f"der{ax}{ax}_11", odd, A, 0, axis)
            worst = max(worst, np.abs(d + _kprime2_d2(k, h) * odd).max() / (k * k * lanes.max()))
    assert worst < tol, worst


def check_cubics_exact(backend, x3d=None, tol=1e-13):
    """_22 (one-sided closures at both ends): every row is at least third-order accurate, so a cubic is differentiated
    exactly (first derivative) and its second derivative is exact as well"""
    n = 41
    length = 3.0
    h = length / (n - 1)
    xs = np.arange(n) * h
    p = 0.7 * xs ** 3 - 1.1 * xs ** 2 + 0.4 * xs + 2.0
    dp = 2.1 * xs ** 2 - 2.2 * xs + 0.4
    ddp = 4.2 * xs - 2.2
    worst = 0.0
    for axis, ax in enumerate("xyz"):
        A = ol.Axis(n, 2, 2, length)
        u, lanes = _field(p, axis)
        e1, _ = _field(dp, axis)
        e2, _ = _field(ddp, axis)
        d1 = _apply(backend, x3d, f"der{ax}_22", u, A, 1, axis)
        d2 = _apply(backend, x3d, f"der{ax}{ax}_22", u, A, 1, axis)
        # exact up to the rounding of the differences: errors scale with max|p| / h and max|p| / h^2
        scale = np.abs(u).max()
        worst = max(worst, np.abs(d1 - e1).max() / (scale / h), np.abs(d2 - e2).max() / (scale / (h * h)))
    assert worst < tol, worst


def check_filter_transfer_function(backend, x3d=None, tol=2e-13):
    """periodic filter (Gaitonde & Visbal 6th-order tridiagonal filter, src/filters.f90:87-93): a Fourier mode is scaled by
    T(k) = [a + b cos(kh) + c cos(2kh) + d cos(3kh)] / (1 + 2 af cos(kh)),
    a = (11 + 10 af)/16, b = (15 + 34 af)/32, c = (-3 + 6 af)/16, d = (1 - 2 af)/32; T(0) = 1, T(pi/h) = 0"""
    h = LEN / N
    xs = np.arange(N) * h
    worst = 0.0
    for af in (0.45, 0.3, -0.2):
        a, b, c, d = (11 + 10 * af) / 16, (15 + 34 * af) / 32, (-3 + 6 * af) / 16, (1 - 2 * af) / 32
        for axis, ax in enumerate("xyz"):
            A = ol.Axis(N, 0, 0, LEN, af=af)
            for m in PERIODIC_MODES + [0, N // 2]:
                k = 2 * np.pi * m / LEN
                T = (a + b * np.cos(k * h) + c * np.cos(2 * k * h) + d * np.cos(3 * k * h)) / (1 + 2 * af * np.cos(k * h))
                u, lanes = _field(np.cos(k * xs), axis)
                f = _apply(backend, x3d, f"fil{ax}_00", u, A, 1, axis)
                worst = max(worst, np.abs(f - T * u).max() / lanes.max())
            assert abs((a + b + c + d) / (1 + 2 * af) - 1) < 1e-15 and abs(a - b + c - d) < 1e-15
    assert worst < tol, worst


def test_oracle_filter_transfer_function():
    check_filter_transfer_function("oracle")


def test_oracle_periodic_modified_wavenumbers():
    check_periodic_modified_wavenumbers("oracle")


def test_oracle_symmetric_modes():
    check_symmetric_modes("oracle")


def test_oracle_cubics_exact():
    check_cubics_exact("oracle")


@pytest.fixture
def x3d():
    from incompact3d_b200 import X3D
    x = X3D(0)
    yield x
    x.close()


@pytest.mark.gpu
def test_product_periodic_modified_wavenumbers(x3d):
    check_periodic_modified_wavenumbers("product", x3d, tol=2e-12)


@pytest.mark.gpu
def test_product_symmetric_modes(x3d):
    check_symmetric_modes("product", x3d, tol=2e-12)


@pytest.mark.gpu
def test_product_cubics_exact(x3d):
    check_cubics_exact("product", x3d, tol=1e-12)


@pytest.mark.gpu
def test_product_filter_transfer_function(x3d):
    check_filter_transfer_function("product", x3d, tol=2e-12)
