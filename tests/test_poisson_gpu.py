"""GPU parity of the spectral Poisson solver (x3d_poisson, C ABI) against the oracle's
restatement of src/poisson.f90 for the four solver variants 000 / 100 / 010 / 11x (bcz=0,1).
FFT engines differ (cuFFT vs the oracle's plain DFT), so the tolerance is 1e-11 relative L-inf."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def oracle_poisson(nn, ncl, lengths):
    L = ol.lib()
    L.x3do_poisson_create.restype = C.c_void_p
    L.x3do_poisson_create.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 3 + [C.c_int] * 2
    L.x3do_poisson_solve.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.x3do_poisson_destroy.argtypes = [C.c_void_p]
    h = L.x3do_poisson_create(nn[0], nn[1], nn[2], (C.c_int * 6)(*ncl), lengths[0], lengths[1], lengths[2], 4, 3)
    assert h, L.x3do_last_error()
    return L, C.c_void_p(h)


CASES = [
    ((16, 12, 20), (0, 0, 0, 0, 0, 0)),   # poisson_000
    ((17, 12, 20), (2, 2, 0, 0, 0, 0)),   # poisson_100
    ((16, 13, 20), (0, 0, 2, 2, 0, 0)),   # poisson_010
    ((17, 13, 20), (1, 1, 1, 1, 0, 0)),   # poisson_11x, bcz=0
    ((17, 13, 21), (1, 1, 1, 1, 1, 1)),   # poisson_11x, bcz=1 (TGV tests/ configuration)
    ((33, 65, 17), (2, 1, 1, 2, 2, 2)),
    ((64, 64, 64), (0, 0, 0, 0, 0, 0)),
    # power-of-two periodic meshes: the hand-written FFT passes (x3d_fft_kernels.cuh): radices 8.8, 8.8.2, 8.8.4, 8.8.8 for the
    # complex lines, half-length transforms 8.4, 8.8, 8.8.2, 8.8.4 for the real ones
    ((128, 256, 64), (0, 0, 0, 0, 0, 0)),
    ((256, 64, 128), (0, 0, 0, 0, 0, 0)),
    ((64, 128, 256), (0, 0, 0, 0, 0, 0)),
    ((512, 64, 64), (0, 0, 0, 0, 0, 0)),
    ((64, 512, 64), (0, 0, 0, 0, 0, 0)),
    ((64, 64, 512), (0, 0, 0, 0, 0, 0)),
]


def test_poisson_000_cufft_path_matches_oracle(monkeypatch):
    """X3D_FFT=0: the library FFT plans + k_spec_000s instead of the hand-written passes"""
    monkeypatch.setenv("X3D_FFT", "0")
    test_poisson_matches_oracle((64, 64, 64), (0, 0, 0, 0, 0, 0))


STRETCHED = [  # (nodes, ncl, istret): matrice_refinement + inversion5_v1/v2 (src/poisson.f90:1814, src/tools.f90:1225)
    ((16, 13, 20), (0, 0, 2, 2, 0, 0), 1), ((16, 13, 20), (0, 0, 2, 2, 0, 0), 2), ((16, 13, 20), (0, 0, 2, 2, 0, 0), 3),
    ((17, 13, 20), (2, 2, 2, 2, 0, 0), 2), ((17, 13, 21), (1, 1, 1, 1, 1, 1), 1), ((17, 17, 21), (2, 2, 2, 2, 2, 2), 3),
    ((64, 65, 32), (0, 0, 2, 2, 0, 0), 2),   # channel-like (BASELINE config #3 at reduced size)
]


@pytest.mark.parametrize("nn,ncl,istret", STRETCHED)
def test_stretched_poisson_matches_oracle(nn, ncl, istret):
    from incompact3d_b200 import X3D, AxisSchemes
    lengths = (2 * np.pi, 2.0, 1.7)
    beta = 0.259065151
    L = ol.lib()
    L.x3do_poisson_create_stretched.restype = C.c_void_p
    L.x3do_poisson_create_stretched.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 3 + [C.c_int] * 3 + [C.c_double]
    L.x3do_poisson_solve.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.x3do_poisson_destroy.argtypes = [C.c_void_p]
    L.x3do_stretching.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    h = L.x3do_poisson_create_stretched(nn[0], nn[1], nn[2], (C.c_int * 6)(*ncl), *lengths, 4, 3, istret, beta)
    assert h, L.x3do_last_error()
    h = C.c_void_p(h)
    x = X3D(0)
    axes = [AxisSchemes(nn[a], ncl[2 * a], ncl[2 * a + 1], lengths[a]) for a in range(3)]
    for a in range(3):
        x.set_deriv_coeffs(a, axes[a].c)
    # the Fortran host passes the alpha that stretching() computed (module mod_stret)
    out8 = np.zeros(8 * nn[1])
    alpha = C.c_double()
    assert L.x3do_stretching(istret, beta, lengths[1], nn[1], axes[1].nm, out8.ctypes.data_as(C.POINTER(C.c_double)), C.byref(alpha)) == 0
    bc = [0 if axes[a].periodic else 1 for a in range(3)]
    x.poisson_init(nn[0], nn[1], nn[2], bc[0], bc[1], bc[2], *lengths, istret=istret, alpha=alpha.value, beta=beta)
    shape = tuple(axes[a].nm for a in range(3))
    rng = np.random.default_rng(5 + sum(nn) + istret)
    rhs = np.asfortranarray(rng.uniform(-1, 1, size=shape))
    ref = rhs.copy(order="F")
    assert L.x3do_poisson_solve(h, ref.ctypes.data_as(C.POINTER(C.c_double))) == 0
    got = rhs.copy(order="F")
    x.poisson(got)
    err = H.rel_linf(got, ref)
    assert err < 1e-10, err
    L.x3do_poisson_destroy(h)
    x.close()


@pytest.mark.parametrize("nn,ncl", CASES)
def test_poisson_matches_oracle(nn, ncl):
    from incompact3d_b200 import X3D, AxisSchemes
    lengths = (2 * np.pi, 3.0, 1.7)
    x = X3D(0)
    axes = [AxisSchemes(nn[a], ncl[2 * a], ncl[2 * a + 1], lengths[a]) for a in range(3)]
    for a in range(3):
        x.set_deriv_coeffs(a, axes[a].c)
    bc = [0 if axes[a].periodic else 1 for a in range(3)]
    x.poisson_init(nn[0], nn[1], nn[2], bc[0], bc[1], bc[2], *lengths)
    shape = tuple(axes[a].nm for a in range(3))
    rng = np.random.default_rng(11 + sum(nn))
    rhs = np.asfortranarray(rng.uniform(-1, 1, size=shape))
    ref = rhs.copy(order="F")
    L, h = oracle_poisson(nn, ncl, lengths)
    assert L.x3do_poisson_solve(h, ref.ctypes.data_as(C.POINTER(C.c_double))) == 0
    got = rhs.copy(order="F")
    x.poisson(got)
    err = H.rel_linf(got, ref)
    assert err < 1e-11, err
    # device-resident call gives the same numbers
    import torch
    d = torch.from_numpy(np.ascontiguousarray(rhs.transpose(2, 1, 0))).cuda()
    x.poisson(d)
    x.sync()
    assert np.array_equal(d.cpu().numpy().transpose(2, 1, 0), got)
    L.x3do_poisson_destroy(h)
    x.close()
