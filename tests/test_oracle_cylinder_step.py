"""The assembled cylinder-wake step of the oracle (SURVEY config #4 at a reduced size: x inflow / convective outflow,
y and z periodic, iibm = 2, AB3).  The reference has no golden output for this case; the pieces are pinned one by one
(tests/test_oracle_cyl_golden.py, test_oracle_ibm_golden.py, test_oracle_intt_golden.py, test_oracle_step_golden.py).
This file guards the assembly: with an empty geometry the immersed-boundary step is the plain step bit for bit, and a
short run past a cylinder stays bounded, keeps its inflow plane and slows the fluid down behind the body.  CPU only."""
import ctypes as C

import numpy as np

import oracle_lib as ol

_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
NN = (33, 32, 8)
LEN = (8.0, 6.0, 2.0)
CEX, CEY, RA = 3.0, 3.0, 0.5
NOBJMAX, NPIF, IZAP = 1, 2, 1


def _solver(iibm_geometry):
    L = ol.lib()
    L.x3do_solver_create_case.restype = C.c_void_p
    L.x3do_solver_create_case.argtypes = [C.c_int] * 3 + [_ip] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double, C.c_int, C.c_double, C.c_double]
    s = L.x3do_solver_create_case(*NN, (C.c_int * 6)(2, 2, 0, 0, 0, 0), *LEN, 300.0, 0.005, 3, 4, 4, 3, 0, 0.0, 5, 4.0, 0.44)
    assert s, L.x3do_last_error()
    s = C.c_void_p(s)
    L.x3do_solver_init_cyl.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.x3do_solver_set_ibm.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.x3do_solver_set_ibm_geometry.argtypes = [C.c_void_p] + [C.c_int] * 4 + [_ip, _dp, _dp, _ip, _ip]
    L.x3do_solver_step.argtypes = [C.c_void_p, C.c_int]
    L.x3do_solver_get_velocity.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.x3do_solver_destroy.argtypes = [C.c_void_p]
    L.x3do_solver_init_cyl(s, 1.0, 1.0)
    if iibm_geometry is not None:
        ep, geo = iibm_geometry
        ubc = np.zeros(3)
        assert L.x3do_solver_set_ibm(s, 2, ep.ctypes.data_as(_dp), ubc.ctypes.data_as(_dp)) == 0
        for axis, (nobj, xi, xf, nip, nfp) in enumerate(geo):
            assert L.x3do_solver_set_ibm_geometry(s, axis, NOBJMAX, NPIF, IZAP, nobj.ctypes.data_as(_ip), xi.ctypes.data_as(_dp),
                                                  xf.ctypes.data_as(_dp), nip.ctypes.data_as(_ip), nfp.ctypes.data_as(_ip)) == 0
    return L, s


def _velocity(L, s):
    u = [np.zeros(NN, order="F") for _ in range(3)]
    L.x3do_solver_get_velocity(s, *[a.ctypes.data_as(_dp) for a in u])
    return u


def _geometry(with_body):
    """what genepsi3d would hand over for a circular cylinder along z (analytic wall positions), or no body at all"""
    nx, ny, nz = NN
    dx, dy = LEN[0] / (nx - 1), LEN[1] / ny
    xs, ys = np.arange(nx) * dx, np.arange(ny) * dy
    ep = np.zeros(NN, order="F")
    geo = []
    shapes = [(ny, nz), (nx, nz), (nx, ny)]
    for axis in range(3):
        na, nb = shapes[axis]
        nobj = np.zeros((na, nb), dtype=np.int32, order="F")
        xi = np.zeros((NOBJMAX, na, nb), order="F"); xf = np.zeros((NOBJMAX, na, nb), order="F")
        nip = np.full((NOBJMAX + 1, na, nb), NPIF, dtype=np.int32, order="F"); nfp = nip.copy(order="F")
        geo.append([nobj, xi, xf, nip, nfp])
    if with_body:
        inside = (xs[:, None] - CEX) ** 2 + (ys[None, :] - CEY) ** 2 <= RA ** 2
        ep[inside, :] = 1.0
        for j in range(ny):
            if abs(ys[j] - CEY) < RA:
                half = np.sqrt(RA ** 2 - (ys[j] - CEY) ** 2)
                geo[0][0][j, :] = 1; geo[0][1][0, j, :] = CEX - half; geo[0][2][0, j, :] = CEX + half
        for i in range(nx):
            if abs(xs[i] - CEX) < RA:
                half = np.sqrt(RA ** 2 - (xs[i] - CEX) ** 2)
                geo[1][0][i, :] = 1; geo[1][1][0, i, :] = CEY - half; geo[1][2][0, i, :] = CEY + half
        for i in range(nx):
            for j in range(ny):
                if inside[i, j]:
                    geo[2][0][i, j] = 1; geo[2][1][0, i, j] = 0.0; geo[2][2][0, i, j] = LEN[2]
    return ep, geo


def test_empty_geometry_is_the_plain_step():
    L, s0 = _solver(None)
    _, s1 = _solver(_geometry(False))
    assert L.x3do_solver_step(s0, 4) == 0 and L.x3do_solver_step(s1, 4) == 0
    for a, b in zip(_velocity(L, s0), _velocity(L, s1)):
        assert np.array_equal(a, b)
    L.x3do_solver_destroy(s0); L.x3do_solver_destroy(s1)


def test_flow_past_a_cylinder_stays_bounded_and_forms_a_wake():
    L, s = _solver(_geometry(True))
    assert L.x3do_solver_step(s, 40) == 0, L.x3do_last_error()
    ux, uy, uz = _velocity(L, s)
    assert np.isfinite(ux).all() and np.isfinite(uy).all() and np.isfinite(uz).all()
    assert np.abs(ux).max() < 2.5 and np.abs(uy).max() < 1.5 and np.abs(uz).max() < 1e-10   # 2-D flow, no noise
    assert np.allclose(ux[0], 1.0)                                  # inflow plane (inflow_noise = 0)
    dx, dy = LEN[0] / (NN[0] - 1), LEN[1] / NN[1]
    i_wake, j_c, j_far = int(round((CEX + 2 * RA) / dx)), int(round(CEY / dy)), 2
    assert ux[i_wake, j_c, 0] < 0.8 * ux[i_wake, j_far, 0]          # velocity deficit right behind the body
    # the flow is deflected around the body: fluid beside it is faster than the free stream
    j_side = int(round((CEY + 2.0 * RA) / dy))
    assert ux[int(round(CEX / dx)), j_side, 0] > 1.0
    L.x3do_solver_destroy(s)
