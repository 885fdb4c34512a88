"""The oracle's time integrators (oracle/solver.cpp Solver::init / Solver::intt) against golden vectors obtained by
executing the reference's own statements (src/variables.f90:1340-1423, src/time_integrators.f90:20-190) -
tests/golden/make_golden_intt.py.  Euler, AB2, AB3 (with their start-up steps) and RK3; bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from test_oracle_tgv import make_solver


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "intt.npz"))


@pytest.mark.parametrize("scheme", [1, 2, 3, 5])
def test_intt_matches_reference_statements(gold, scheme):
    dt = float(gold["meta/dt"])
    L, s = make_solver(n=8, ncl=(0,) * 6, length=2 * np.pi, dt=dt, itimescheme=scheme)
    dp = C.POINTER(C.c_double)
    L.x3do_solver_intt.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, dp, dp]
    L.x3do_solver_time_coefficients.argtypes = [C.c_void_p, dp]
    co = np.zeros(14)
    L.x3do_solver_time_coefficients(s, co.ctypes.data_as(dp))
    for q, nm in enumerate(("adt", "bdt", "cdt", "gdt")):
        ref = gold[f"s{scheme}/{nm}"][:3].copy()
        if nm == "gdt" and scheme != 5:
            ref[2] = 0.0     # the reference also sets gdt(3) = gdt(1) for single-stage schemes; never read by the time step
        assert np.array_equal(co[3 * q:3 * q + 3], ref), (nm, co[3 * q:3 * q + 3], ref)
    ntime, iadv = (int(v) for v in gold[f"s{scheme}/ntime_iadvance"])
    assert (int(co[12]), int(co[13])) == (ntime, iadv)
    var = gold[f"s{scheme}/var0"].copy(order="F")
    n = var.size
    dvar = np.zeros(var.shape + (ntime,), order="F")
    for c in range(int(gold[f"s{scheme}/ncalls"])):
        itime, itr = (int(v) for v in gold[f"s{scheme}/c{c}/itime_itr"])
        dvar[..., 0] = gold[f"s{scheme}/c{c}/rhs"]
        assert L.x3do_solver_intt(s, itime, itr, n, var.ctypes.data_as(dp), dvar.ctypes.data_as(dp)) == 0
        assert np.array_equal(var, gold[f"s{scheme}/c{c}/var"]), (scheme, c)
        if ntime > 1:
            # what later calls read: the stored right-hand sides
            assert np.array_equal(dvar[..., 1:], gold[f"s{scheme}/c{c}/dvar"][..., 1:]), (scheme, c)
    L.x3do_solver_destroy(s)
