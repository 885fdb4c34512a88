"""Oracle vs tests/golden/step.npz -- outputs of the REFERENCE's own statements of pre_correc and the closing part of
gradp (src/navier.f90:502-789, 439-496) and of channel_cfr (src/Case-Channel.f90:220-261), executed from the Fortran
text by tests/golden/make_golden_step.py.  Pins the wall / forcing pieces of the channel-flow step.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

_dp = C.POINTER(C.c_double)
DPD = ["dpdyx1", "dpdzx1", "dpdyxn", "dpdzxn", "dpdxy1", "dpdzy1", "dpdxyn", "dpdzyn", "dpdxz1", "dpdyz1", "dpdxzn", "dpdyzn"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/step.npz")


def _solver(L, nn, ncl):
    L.x3do_solver_create_case.restype = C.c_void_p
    L.x3do_solver_create_case.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double, C.c_int,
                                                                                                                 C.c_double, C.c_double]
    s = L.x3do_solver_create_case(*nn, (C.c_int * 6)(*ncl), 2.0, 2.0, 2.0, 100.0, 0.01, 5, 4, 4, 3, 0, 0.0, 0, 4.0, 0.44)
    assert s, L.x3do_last_error()
    for f in ("x3do_solver_set_wall_gradient", "x3do_solver_get_wall_gradient"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int, _dp]
    L.x3do_solver_pre_correc.argtypes = [C.c_void_p, C.c_int, _dp]
    L.x3do_solver_capture_wall_gradients.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
    L.x3do_solver_set_velocity.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.x3do_solver_get_velocity.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.x3do_solver_destroy.argtypes = [C.c_void_p]
    return C.c_void_p(s)


def _p(a):
    return a.ctypes.data_as(_dp)


@pytest.mark.parametrize("tag", ["y22", "x22z22y11", "x21y12z11"])
def test_pre_correc_and_gradp_capture(gold, tag):
    L = ol.lib()
    ncl = [int(v) for v in gold[f"pre_correc/{tag}/ncl"]]
    u = [np.asfortranarray(gold[f"pre_correc/{tag}/in/{n}"]).copy(order="F") for n in ("ux", "uy", "uz")]
    nn = u[0].shape
    gdt = np.ascontiguousarray(gold["meta/gdt"])
    itr = int(gold["meta/itr"])
    s = _solver(L, nn, ncl)
    L.x3do_solver_set_velocity(s, *[_p(a) for a in u])
    for q, nm in enumerate(DPD):
        L.x3do_solver_set_wall_gradient(s, q, _p(np.asfortranarray(gold[f"pre_correc/{tag}/in/{nm}"]).copy(order="F")))
    assert L.x3do_solver_pre_correc(s, itr, _p(gdt)) == 0, L.x3do_last_error()
    got = [np.zeros(nn, order="F") for _ in range(3)]
    L.x3do_solver_get_velocity(s, *[_p(a) for a in got])
    for a, nm in zip(got, ("ux", "uy", "uz")):
        assert np.array_equal(a, gold[f"pre_correc/{tag}/out/{nm}"]), (tag, nm)
    # the wall gradients are scaled in place only on Dirichlet faces
    for q, nm in enumerate(DPD):
        ref = gold[f"pre_correc/{tag}/out/{nm}"]
        buf = np.zeros(ref.shape, order="F")
        L.x3do_solver_get_wall_gradient(s, q, _p(buf))
        assert np.array_equal(buf, ref), (tag, nm)
    # gradp: capture of the wall gradients from px1, py1, pz1
    p3 = [np.asfortranarray(gold[f"gradp/{tag}/in/{n}"]).copy(order="F") for n in ("px1", "py1", "pz1")]
    assert L.x3do_solver_capture_wall_gradients(s, itr, _p(gdt), *[_p(a) for a in p3]) == 0
    for q, nm in enumerate(DPD):
        ref = gold[f"gradp/{tag}/out/{nm}"]
        buf = np.zeros(ref.shape, order="F")
        L.x3do_solver_get_wall_gradient(s, q, _p(buf))
        assert np.array_equal(buf, ref), (tag, nm)
    L.x3do_solver_destroy(s)


@pytest.mark.parametrize("tag", ["uniform", "stretched"])
def test_channel_cfr(gold, tag):
    L = ol.lib()
    L.x3do_channel_cfr.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_double]
    u = np.asfortranarray(gold[f"cfr/{tag}/in"]).copy(order="F")
    ppy = np.ascontiguousarray(gold[f"cfr/{tag}/ppy"])
    dy, yly = (float(v) for v in gold[f"cfr/{tag}/dy_yly"])
    L.x3do_channel_cfr(_p(u), *u.shape, _p(ppy), dy, yly, 2.0 / 3.0)
    assert np.abs(u - gold[f"cfr/{tag}/out"]).max() < 1e-15
