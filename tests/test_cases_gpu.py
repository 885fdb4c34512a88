"""GPU parity of the case glue of the device time step against the oracle's assembled steps:

 * BASELINE config #4, cylinder wake: inflow + convective outflow planes (src/Case-Cylinder-wake.f90:100-203), pre_correc
   with the wall-velocity planes (src/navier.f90:564-595), iibm = 2 (lagpol* in front of every derivative of
   momentum_rhs_eq, src/derive.f90:23-24; (1 - ep1) u + ep1 ubc in divergence, src/navier.f90:285-293), AB3 -- at
   193x64x16 and at the full 769x256x32;
 * BASELINE config #3, channel at its full size 256x129x128 (istret = 2, isecondder = 5), and the channel's
   momentum_forcing_channel (src/Case-Channel.f90:396-420): spin-up rotation and constant pressure gradient."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
NOBJMAX, NPIF, IZAP = 1, 2, 1


def _oracle(nn, ncl, lens, re, dt, scheme, second, istret, beta, itype):
    L = ol.lib()
    L.x3do_solver_create_case.restype = C.c_void_p
    L.x3do_solver_create_case.argtypes = [C.c_int] * 3 + [_ip] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double, C.c_int, C.c_double, C.c_double]
    s = L.x3do_solver_create_case(*nn, (C.c_int * 6)(*ncl), *lens, re, dt, scheme, 4, second, 3, istret, beta, itype, 4.0, 0.44)
    assert s, L.x3do_last_error()
    L.x3do_solver_get_velocity.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.x3do_solver_step.argtypes = [C.c_void_p, C.c_int]
    L.x3do_solver_destroy.argtypes = [C.c_void_p]
    return L, C.c_void_p(s)


def _velocity(L, s, nn):
    u = [np.zeros(nn, order="F") for _ in range(3)]
    L.x3do_solver_get_velocity(s, *[a.ctypes.data_as(_dp) for a in u])
    return u


def cylinder_geometry(nn, lens, cex, cey, ra):
    """what genepsi3d hands over for a circular cylinder along z (analytic wall positions): ep1 and, per direction,
    nobj / xi / xf / nxipif / nxfpif (src/module_param.f90:546-556)"""
    nx, ny, nz = nn
    dx, dy = lens[0] / (nx - 1), lens[1] / ny
    xs, ys = np.arange(nx) * dx, np.arange(ny) * dy
    ep = np.zeros(nn, order="F")
    geo = []
    for na, nb in ((ny, nz), (nx, nz), (nx, ny)):
        nobj = np.zeros((na, nb), dtype=np.int32, order="F")
        xi = np.zeros((NOBJMAX, na, nb), order="F"); xf = np.zeros((NOBJMAX, na, nb), order="F")
        nip = np.full((NOBJMAX + 1, na, nb), NPIF, dtype=np.int32, order="F"); nfp = nip.copy(order="F")
        geo.append([nobj, xi, xf, nip, nfp])
    inside = (xs[:, None] - cex) ** 2 + (ys[None, :] - cey) ** 2 <= ra ** 2
    ep[inside, :] = 1.0
    for j in range(ny):
        if abs(ys[j] - cey) < ra:
            half = np.sqrt(ra ** 2 - (ys[j] - cey) ** 2)
            geo[0][0][j, :] = 1; geo[0][1][0, j, :] = cex - half; geo[0][2][0, j, :] = cex + half
    for i in range(nx):
        if abs(xs[i] - cex) < ra:
            half = np.sqrt(ra ** 2 - (xs[i] - cex) ** 2)
            geo[1][0][i, :] = 1; geo[1][1][0, i, :] = cey - half; geo[1][2][0, i, :] = cey + half
    for i in range(nx):
        for j in range(ny):
            if inside[i, j]:
                geo[2][0][i, j] = 1; geo[2][1][0, i, j] = 0.0; geo[2][2][0, i, j] = lens[2]
    return ep, geo, dx, dy


@pytest.mark.parametrize("nn,noise", [((193, 64, 16), 0.0), ((193, 64, 16), 0.05), ((769, 256, 32), 0.0)])
def test_cylinder_step_matches_oracle(nn, noise):
    from incompact3d_b200 import X3D
    lens = (20.0, 12.0, 6.0)
    cex, cey, ra = 5.0, 6.0, 0.5
    ncl = (2, 2, 0, 0, 0, 0)
    re, dt, nsteps = 300.0, 0.0025, 5
    L, s = _oracle(nn, ncl, lens, re, dt, 3, 4, 0, 0.0, 5)
    L.x3do_solver_init_cyl.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.x3do_solver_set_ibm.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    L.x3do_solver_set_ibm_geometry.argtypes = [C.c_void_p] + [C.c_int] * 4 + [_ip, _dp, _dp, _ip, _ip]
    L.x3do_solver_set_inflow_noise.argtypes = [C.c_void_p, C.c_double, _dp, _dp, _dp]
    L.x3do_solver_init_cyl(s, 1.0, 1.0)
    ep, geo, dx, dy = cylinder_geometry(nn, lens, cex, cey, ra)
    ubc = np.zeros(3)
    assert L.x3do_solver_set_ibm(s, 2, ep.ctypes.data_as(_dp), ubc.ctypes.data_as(_dp)) == 0
    for axis, (nobj, xi, xf, nip, nfp) in enumerate(geo):
        assert L.x3do_solver_set_ibm_geometry(s, axis, NOBJMAX, NPIF, IZAP, nobj.ctypes.data_as(_ip), xi.ctypes.data_as(_dp),
                                              xf.ctypes.data_as(_dp), nip.ctypes.data_as(_ip), nfp.ctypes.data_as(_ip)) == 0
    rng = np.random.default_rng(7)
    planes = [np.asfortranarray(rng.uniform(-0.5, 0.5, size=(nn[1], nn[2]))) for _ in range(3)]
    if noise:
        L.x3do_solver_set_inflow_noise(s, noise, *[a.ctypes.data_as(_dp) for a in planes])

    x = X3D(0)
    x.solver_init(*nn, ncl=ncl, xlx=lens[0], yly=lens[1], zlz=lens[2], re=re, dt=dt, itimescheme=3, itype=5)
    x.solver_set_case(u1=1.0, u2=1.0, inflow_noise=noise, iibm=2, ubc=(0.0, 0.0, 0.0))
    x.solver_set_ibm_mask(ep)
    if noise:
        x.solver_set_inflow_noise(*planes)
    dz = lens[2] / nn[2]
    for axis, (nobj, xi, xf, nip, nfp) in enumerate(geo):
        x.set_ibm_geometry(axis, NOBJMAX, NPIF, IZAP, nobj, xi, xf, nip, nfp, (dx, dy, dz)[axis], lens[axis],
                           coords=np.arange(nn[1]) * dy if axis == 1 else None)
    x.solver_init_cyl()
    for it in range(nsteps):
        x.solver_step(1)
        assert L.x3do_solver_step(s, 1) == 0, L.x3do_last_error()
        got = x.solver_get_velocity()
        ref = _velocity(L, s, nn)
        scale = max(np.abs(r).max() for r in ref)
        err = max(np.abs(a - b).max() for a, b in zip(got, ref)) / scale
        assert err < 1e-10, (it, err)
    # the body slows the fluid down and the inflow plane holds u1 (+ noise)
    assert np.allclose(got[0][0], 1.0 + noise * planes[0] if noise else 1.0)
    L.x3do_solver_destroy(s)
    x.close()


def test_channel_full_size_matches_oracle():
    """BASELINE config #3 at its real size 256 x 129 x 128: stretched walls (istret = 2), hyperviscous second derivative,
    stretched Poisson (poisson_010 + inversion5_v1), constant flow rate; 4 RK3 steps"""
    from incompact3d_b200 import X3D
    nn, ncl, lens, beta = (256, 129, 128), (0, 0, 2, 2, 0, 0), (8.0, 2.0, 4.0), 0.259065151
    L, s = _oracle(nn, ncl, lens, 4200.0, 0.005, 5, 5, 2, beta, 3)
    L.x3do_solver_init_channel.argtypes = [C.c_void_p]
    L.x3do_solver_init_channel(s)
    x = X3D(0)
    x.solver_init(*nn, ncl=ncl, xlx=lens[0], yly=lens[1], zlz=lens[2], re=4200.0, dt=0.005, isecondder=5, istret=2, beta=beta, itype=3)
    x.solver_init_channel()
    for it in range(4):
        x.solver_step(1)
        assert L.x3do_solver_step(s, 1) == 0, L.x3do_last_error()
        got = x.solver_get_velocity()
        ref = _velocity(L, s, nn)
        scale = max(np.abs(r).max() for r in ref)
        err = max(np.abs(a - b).max() for a, b in zip(got, ref)) / scale
        assert err < 1e-10, (it, err)
    dmax, _ = x.solver_divergence()
    assert abs(dmax) < 1e-9
    L.x3do_solver_destroy(s)
    x.close()


@pytest.mark.parametrize("cpg,wrot", [(0, 0.12), (1, 0.0), (1, 0.12)])
def test_channel_forcing_matches_oracle(cpg, wrot):
    """momentum_forcing_channel (src/Case-Channel.f90:396-420): spin-up rotation for itime < spinup_time (here it ends
    after the second step), constant pressure gradient with the re_cent viscosity (src/parameters.f90:303-311)"""
    from incompact3d_b200 import X3D
    nn, ncl, lens, beta = (32, 33, 16), (0, 0, 2, 2, 0, 0), (8.0, 2.0, 3.0), 0.259065151
    re = 180.0 if cpg else 4200.0
    L, s = _oracle(nn, ncl, lens, re, 0.002, 5, 5, 2, beta, 3)
    L.x3do_solver_init_channel.argtypes = [C.c_void_p]
    L.x3do_solver_set_channel_forcing.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
    L.x3do_solver_set_channel_forcing(s, cpg, wrot, 3, 1)
    L.x3do_solver_init_channel(s)
    x = X3D(0)
    x.solver_init(*nn, ncl=ncl, xlx=lens[0], yly=lens[1], zlz=lens[2], re=re, dt=0.002, isecondder=5, istret=2, beta=beta, itype=3)
    x.solver_set_case(cpg=cpg, wrotation=wrot, spinup_time=3, iin=1)
    x.solver_init_channel()
    for it in range(4):
        x.solver_step(1)
        assert L.x3do_solver_step(s, 1) == 0, L.x3do_last_error()
        got = x.solver_get_velocity()
        ref = _velocity(L, s, nn)
        scale = max(np.abs(r).max() for r in ref)
        err = max(np.abs(a - b).max() for a, b in zip(got, ref)) / scale
        assert err < 1e-10, (it, cpg, wrot, err)
    L.x3do_solver_destroy(s)
    x.close()


@pytest.mark.parametrize("ncl", [(0,) * 6, (1,) * 6, (2, 2, 0, 0, 0, 0), (0, 0, 2, 2, 0, 0)])
@pytest.mark.parametrize("ifilter", [1, 2, 3])
def test_apply_spatial_filter_matches_oracle(ncl, ifilter):
    """apply_spatial_filter (src/tools.f90:600-675) on the solver's velocity against the oracle's filx / fily / filz
    composed with the npaire pairing of the reference"""
    import helpers as H
    from incompact3d_b200 import X3D
    nn = (40 + (ncl[0] != 0), 36 + (ncl[2] != 0), 32 + (ncl[4] != 0))
    lens = (2 * np.pi, 2.0, 3.0)
    af = 0.45
    rng = np.random.default_rng(3)
    vel = [np.asfortranarray(rng.uniform(-1, 1, size=nn)) for _ in range(3)]
    x = X3D(0)
    x.solver_init(*nn, ncl=ncl, xlx=lens[0], yly=lens[1], zlz=lens[2], re=1000.0, dt=0.001, itype=0)
    x.solver_set_velocity(*vel)
    x.solver_apply_spatial_filter(ifilter, af)
    got = x.solver_get_velocity()
    axes = [ol.Axis(nn[a], ncl[2 * a], ncl[2 * a + 1], lens[a], af=af) for a in range(3)]
    ref = [v.copy(order="F") for v in vel]
    for a, on in ((0, ifilter in (1, 2)), (1, ifilter in (1, 3)), (2, ifilter in (1, 2))):
        if not on:
            continue
        name = f"fil{'xyz'[a]}_{ncl[2 * a]}{ncl[2 * a + 1]}"
        ref = [H.oracle_op(name, ref[c], axes[a], 0 if c == a else 1) for c in range(3)]
    for c in range(3):
        assert H.rel_linf(got[c], ref[c]) < 1e-12, (c, H.rel_linf(got[c], ref[c]))
    x.close()
