"""Oracle vs tests/golden/cyl.npz -- outputs of the REFERENCE's own statements of the inflow / convective outflow planes
(src/Case-Cylinder-wake.f90:100-203), of body / corgp_IBM (src/ibm.f90:14-80) and of pre_correc with non-zero wall
velocities, without (itype = cylinder) and with (itype = channel) the inflow / outflow flow-rate correction
(src/navier.f90:534-595), executed from the Fortran text by tests/golden/make_golden_cyl.py.  These are the pieces of
the cylinder-wake step (second half) that the oracle restates ahead of the library.  CPU only, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

_dp = C.POINTER(C.c_double)
WALLS = ["bxx1", "bxy1", "bxz1", "bxxn", "bxyn", "bxzn", "byx1", "byy1", "byz1", "byxn", "byyn", "byzn",
         "bzx1", "bzy1", "bzz1", "bzxn", "bzyn", "bzzn"]
DPD = ["dpdyx1", "dpdzx1", "dpdyxn", "dpdzxn", "dpdxy1", "dpdzy1", "dpdxyn", "dpdzyn", "dpdxz1", "dpdyz1", "dpdxzn", "dpdyzn"]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/cyl.npz")


def _p(a):
    return a.ctypes.data_as(_dp)


def _solver(L, nn, ncl, itype, xlx):
    L.x3do_solver_create_case.restype = C.c_void_p
    L.x3do_solver_create_case.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double, C.c_int,
                                                                                                                 C.c_double, C.c_double]
    s = L.x3do_solver_create_case(*nn, (C.c_int * 6)(*ncl), xlx, 2.0, 2.0, 100.0, 0.01, 5, 4, 4, 3, 0, 0.0, itype, 4.0, 0.44)
    assert s, L.x3do_last_error()
    for f in ("x3do_solver_set_wall_gradient", "x3do_solver_get_wall_gradient", "x3do_solver_set_wall_velocity", "x3do_solver_get_wall_velocity"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int, _dp]
    L.x3do_solver_pre_correc.argtypes = [C.c_void_p, C.c_int, _dp]
    L.x3do_solver_inflow_outflow.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]
    L.x3do_solver_set_velocity.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.x3do_solver_get_velocity.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.x3do_solver_destroy.argtypes = [C.c_void_p]
    return C.c_void_p(s)


@pytest.mark.parametrize("tag", ["u1_0", "u1_1", "u1_2", "u1_07"])
def test_inflow_outflow_planes(gold, tag):
    L = ol.lib()
    u = [np.asfortranarray(gold[f"outflow/in/{n}"]).copy(order="F") for n in ("ux", "uy", "uz")]
    nn = u[0].shape
    dx = float(gold["meta/dx"])
    # x is a Dirichlet direction: dx = xlx / (nx - 1)
    s = _solver(L, nn, (2, 2, 0, 0, 0, 0), 5, dx * (nn[0] - 1))
    L.x3do_solver_set_velocity(s, *[_p(a) for a in u])
    u1, u2 = (float(v) for v in gold[f"outflow/{tag}/u1_u2"])
    noise = float(gold["inflow/u1_noise"][1])
    planes = [np.asfortranarray(gold[f"inflow/in/{n}"]).copy(order="F") for n in ("bxo", "byo", "bzo")]
    gdt = np.ascontiguousarray(gold["meta/gdt"])
    assert L.x3do_solver_inflow_outflow(s, int(gold["meta/itr"]), _p(gdt), u1, u2, noise, *[_p(a) for a in planes]) == 0
    for q, nm in enumerate(WALLS[:6]):
        buf = np.zeros((nn[1], nn[2]), order="F")
        L.x3do_solver_get_wall_velocity(s, q, _p(buf))
        if q < 3:
            ref = u1 + planes[0] * noise if q == 0 else planes[q] * noise    # same expressions as the reference's inflow
            if u1 == 1.0:
                assert np.array_equal(buf, gold[f"inflow/out/{nm}"]), nm
            assert np.array_equal(buf, ref), nm
        else:
            assert np.array_equal(buf, gold[f"outflow/{tag}/{nm}"]), (tag, nm)
    L.x3do_solver_destroy(s)


def test_body_and_corgp(gold):
    L = ol.lib()
    L.x3do_ibm_body.argtypes = [_dp, _dp, _dp, _dp, C.c_longlong]
    L.x3do_ibm_corgp.argtypes = [_dp] * 6 + [C.c_longlong, C.c_int]
    u0 = [np.asfortranarray(gold[f"outflow/in/{n}"]) for n in ("ux", "uy", "uz")]
    ep = np.asfortranarray(gold["body/in/ep"]).copy(order="F")
    u = [a.copy(order="F") for a in u0]
    L.x3do_ibm_body(*[_p(a) for a in u], _p(ep), u[0].size)
    for a, nm in zip(u, ("ux", "uy", "uz")):
        assert np.array_equal(a, gold[f"body/out/{nm}"])
    p3 = [np.asfortranarray(gold[f"corgp/in/{n}"]).copy(order="F") for n in ("px", "py", "pz")]
    for nlock in (1, 2):
        u = [a.copy(order="F") for a in u0]
        L.x3do_ibm_corgp(*[_p(a) for a in u], *[_p(a) for a in p3], u[0].size, nlock)
        for a, nm in zip(u, ("ux", "uy", "uz")):
            assert np.array_equal(a, gold[f"corgp/nlock{nlock}/{nm}"])


@pytest.mark.parametrize("tag,itype", [("cyl", 5), ("channel", 3)])
def test_pre_correc_with_wall_velocities(gold, tag, itype):
    L = ol.lib()
    u = [np.asfortranarray(gold[f"outflow/in/{n}"]).copy(order="F") for n in ("ux", "uy", "uz")]
    nn = u[0].shape
    s = _solver(L, nn, (2, 2, 0, 0, 0, 0), itype, 2.0)
    L.x3do_solver_set_velocity(s, *[_p(a) for a in u])
    for q, nm in enumerate(WALLS):
        L.x3do_solver_set_wall_velocity(s, q, _p(np.asfortranarray(gold[f"pre_correc/{tag}/in/{nm}"]).copy(order="F")))
    for q, nm in enumerate(DPD):
        L.x3do_solver_set_wall_gradient(s, q, _p(np.asfortranarray(gold[f"pre_correc/{tag}/in/{nm}"]).copy(order="F")))
    gdt = np.ascontiguousarray(gold["meta/gdt"])
    assert L.x3do_solver_pre_correc(s, int(gold["meta/itr"]), _p(gdt)) == 0, L.x3do_last_error()
    got = [np.zeros(nn, order="F") for _ in range(3)]
    L.x3do_solver_get_velocity(s, *[_p(a) for a in got])
    for a, nm in zip(got, ("ux", "uy", "uz")):
        assert np.array_equal(a, gold[f"pre_correc/{tag}/out/{nm}"]), (tag, nm)
    buf = np.zeros((nn[1], nn[2]), order="F")
    L.x3do_solver_get_wall_velocity(s, 3, _p(buf))
    assert np.array_equal(buf, gold[f"pre_correc/{tag}/out/bxxn"])
    changed = not np.array_equal(buf, gold[f"pre_correc/{tag}/in/bxxn"])
    assert changed == (tag == "channel")     # the flow-rate correction is not applied to the cylinder case
    L.x3do_solver_destroy(s)


@pytest.mark.parametrize("tag", ["cpg", "rot", "rot_over", "both"])
def test_channel_momentum_forcing(gold, tag):
    """momentum_forcing_channel (src/Case-Channel.f90:396-420): constant pressure gradient, spin-up rotation"""
    L = ol.lib()
    u = [np.asfortranarray(gold[f"outflow/in/{n}"]).copy(order="F") for n in ("ux", "uy", "uz")]
    nn = u[0].shape
    s = _solver(L, nn, (0, 0, 2, 2, 0, 0), 3, 2.0)
    L.x3do_solver_set_velocity(s, *[_p(a) for a in u])
    L.x3do_solver_momentum_forcing.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]
    cpg, fcpg, itime, spin, wrot = (float(v) for v in gold[f"forcing/{tag}/par"])
    d = [np.asfortranarray(gold[f"forcing/in/{n}"]).copy(order="F") for n in ("dux", "duy", "duz")]
    assert L.x3do_solver_momentum_forcing(s, int(itime), int(cpg), fcpg, wrot, int(spin), 1, *[_p(a) for a in d]) == 0
    for a, nm in zip(d, ("dux", "duy", "duz")):
        assert np.array_equal(a, gold[f"forcing/{tag}/{nm}"]), (tag, nm)
    L.x3do_solver_destroy(s)
