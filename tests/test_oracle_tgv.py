"""The oracle's time step vs the reference's own golden file
(tests/TGV-Taylor-Green-vortex/reference_time_evol.dat, real Fortran/MPI output, copied to
tests/golden/tgv_reference_time_evol.dat): TGV 65^3 free-slip, Re=1600, dt=0.005, RK3.
BASELINE config #1: rows 1-10 = 100 steps; E_k and enstrophy within 1e-9 (the file holds 12
digits).  This pins der*_11, der**_11, the 12 staggered operators (non-periodic branches),
poisson_11x(bcz=1), waves/abxyz, RK3 and pre_correc.  CPU only."""
import ctypes as C
import os

import numpy as np

import oracle_lib as ol

PI_IN = 3.14159265358979  # xlx in reference_input.i3d:21-23


def make_solver(n=65, ncl=(1, 1, 1, 1, 1, 1), length=PI_IN, re=1600.0, dt=0.005, itimescheme=5):
    L = ol.lib()
    L.x3do_solver_create.restype = C.c_void_p
    L.x3do_solver_create.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double]
    arr = (C.c_int * 6)(*ncl)
    nn = (n, n, n) if isinstance(n, int) else n
    s = L.x3do_solver_create(nn[0], nn[1], nn[2], arr, length, length, length, re, dt, itimescheme, 4, 4, 3, 0, 0.0)
    assert s, L.x3do_last_error()
    return L, C.c_void_p(s)


def test_tgv_65_matches_reference_golden(golden_dir):
    ref = np.loadtxt(os.path.join(golden_dir, "tgv_reference_time_evol.dat"))
    L, s = make_solver()
    L.x3do_solver_init_tgv(s)
    out = (C.c_double * 4)()
    L.x3do_solver_postprocess_tgv(s, out)
    assert abs(out[0] - 0.125) < 1e-4 and abs(out[3] - 0.375) < 1e-3  # TGV t=0: E_k = 1/8, enstrophy = 3/8
    worst = 0.0
    for row in range(10):
        assert L.x3do_solver_step(s, 10) == 0, L.x3do_last_error()
        L.x3do_solver_postprocess_tgv(s, out)
        got = np.array(out[:])
        err = np.abs(got / ref[row, 1:] - 1.0)
        worst = max(worst, err.max())
        assert err[0] < 1e-9 and err[3] < 1e-9, (row, err)   # E_k, enstrophy (BASELINE tolerance)
        assert err[1] < 1e-9 and err[2] < 1e-9, (row, err)   # eps, eps2
        tmax, tmoy = C.c_double(), C.c_double()
        L.x3do_solver_divergence(s, C.byref(tmax), C.byref(tmoy))
        assert abs(tmax.value) < 1e-11 and tmoy.value < 1e-12  # DIV U max / mean at machine level
    print("worst relative deviation from the reference golden file:", worst)
    L.x3do_solver_destroy(s)


def test_adams_bashforth_consistent_with_rk3():
    """AB2 / AB3 (time_integrators.f90:75-100) are restated without golden data: check them against the pinned RK3
    path - same flow, small time step, the kinetic-energy histories agree to the schemes' truncation error."""
    out = {}
    for scheme in (2, 3, 5):
        L, s = make_solver(n=33, dt=0.001, itimescheme=scheme)
        L.x3do_solver_init_tgv(s)
        assert L.x3do_solver_step(s, 12) == 0
        d = (C.c_double * 4)()
        L.x3do_solver_postprocess_tgv(s, d)
        out[scheme] = np.array(d[:])
        L.x3do_solver_destroy(s)
    assert np.abs(out[3] / out[5] - 1).max() < 2e-6, out
    assert np.abs(out[2] / out[5] - 1).max() < 2e-5, out
    assert not np.array_equal(out[2], out[3])
