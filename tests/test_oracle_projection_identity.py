"""div(grad(poisson(rhs))) == rhs for rhs = div(grad(q)): the assembled spectral solves (poisson_000 / 100 / 010 / 11x, uniform meshes) are the
exact inverses of the discrete divergence-of-gradient built from the twelve staggered operators.  The
reference has no golden output for the assembled poisson_000/100/010; this identity pins them (and the modified
wavenumbers / transfer functions of `waves`) against the operators, which are pinned by tests/golden/operators.npz.
Oracle only, CPU."""
import ctypes as C

import numpy as np
import pytest

from test_oracle_tgv import make_solver

CASES = [
    ((16, 12, 20), (0, 0, 0, 0, 0, 0)),   # poisson_000
    ((17, 12, 20), (1, 1, 0, 0, 0, 0)),   # poisson_100
    ((17, 12, 20), (2, 2, 0, 0, 0, 0)),   # poisson_100, Dirichlet velocity walls (same pressure operator)
    ((16, 13, 20), (0, 0, 1, 1, 0, 0)),   # poisson_010
    ((16, 13, 20), (0, 0, 2, 2, 0, 0)),
    ((17, 13, 20), (1, 1, 1, 1, 0, 0)),   # poisson_11x, bcz = 0
    ((17, 13, 21), (1, 1, 1, 1, 1, 1)),   # poisson_11x, bcz = 1
    ((17, 13, 21), (2, 2, 1, 1, 2, 2)),
]


@pytest.mark.parametrize("nn,ncl", CASES)
def test_divergence_of_gradient_of_solution_is_rhs(nn, ncl):
    L, s = make_solver(n=nn, ncl=ncl, length=2 * np.pi, dt=0.001)
    dp = C.POINTER(C.c_double)
    L.x3do_solver_pdims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.x3do_solver_poisson.argtypes = [C.c_void_p, dp]
    L.x3do_solver_gradp.argtypes = [C.c_void_p, dp, dp, dp, dp]
    L.x3do_solver_divergence_of.argtypes = [C.c_void_p, dp, C.c_int]
    L.x3do_solver_set_velocity.argtypes = [C.c_void_p, dp, dp, dp]
    d3 = (C.c_int * 3)()
    L.x3do_solver_pdims(s, d3)
    pd = tuple(d3)
    rng = np.random.default_rng(17 + sum(nn) + sum(ncl))

    def div_grad(q):
        px, py, pz = (np.zeros(nn, order="F") for _ in range(3))
        assert L.x3do_solver_gradp(s, px.ctypes.data_as(dp), py.ctypes.data_as(dp), pz.ctypes.data_as(dp), q.ctypes.data_as(dp)) == 0
        L.x3do_solver_set_velocity(s, px.ctypes.data_as(dp), py.ctypes.data_as(dp), pz.ctypes.data_as(dp))
        out = np.zeros(pd, order="F")
        assert L.x3do_solver_divergence_of(s, out.ctypes.data_as(dp), 2) == 0
        return out

    # a right-hand side in the range of the discrete operator (its null space holds more than the mean: the staggered
    # interpolators annihilate the Nyquist modes, which the solver zeroes, src/poisson.f90:366)
    rhs = div_grad(np.asfortranarray(rng.uniform(-1, 1, pd)))
    p = rhs.copy(order="F")
    assert L.x3do_solver_poisson(s, p.ctypes.data_as(dp)) == 0
    back = div_grad(p)
    err = np.abs(back - rhs).max() / np.abs(rhs).max()
    assert err < 1e-11, err
    L.x3do_solver_destroy(s)


STRETCHED = [  # (nodes, ncl, istret): matrice_refinement + inversion5_v1/v2 (src/poisson.f90:1814, src/tools.f90:1225)
    ((16, 13, 20), (0, 0, 2, 2, 0, 0), 1), ((16, 13, 20), (0, 0, 2, 2, 0, 0), 2), ((17, 13, 20), (2, 2, 2, 2, 0, 0), 2),
    ((17, 13, 21), (1, 1, 1, 1, 1, 1), 1), ((16, 33, 12), (0, 0, 1, 1, 0, 0), 2), ((16, 13, 20), (0, 0, 2, 2, 0, 0), 3),
]


@pytest.mark.parametrize("nn,ncl,istret", STRETCHED)
def test_stretched_solves_invert_the_stretched_operators(nn, ncl, istret):
    """The same identity with the y mesh stretched: the pentadiagonal systems of matrice_refinement are the exact
    spectral image of div(grad) with the ppy / ppyi metric factors for istret = 1 and 2.  For istret = 3 the reference's
    construction is not an exact inverse (the residual below is a property of the algorithm, present in the oracle and
    in the library alike); the test records its size."""
    import oracle_lib as ol
    L = ol.lib()
    L.x3do_solver_create.restype = C.c_void_p
    L.x3do_solver_create.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 5 + [C.c_int] * 5 + [C.c_double]
    s = L.x3do_solver_create(nn[0], nn[1], nn[2], (C.c_int * 6)(*ncl), 2 * np.pi, 2.0, 2 * np.pi, 1000.0, 0.001, 5, 4, 4, 3,
                             istret, 0.259065151)
    assert s, L.x3do_last_error()
    s = C.c_void_p(s)
    dp = C.POINTER(C.c_double)
    L.x3do_solver_pdims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.x3do_solver_poisson.argtypes = [C.c_void_p, dp]
    L.x3do_solver_gradp.argtypes = [C.c_void_p, dp, dp, dp, dp]
    L.x3do_solver_divergence_of.argtypes = [C.c_void_p, dp, C.c_int]
    L.x3do_solver_set_velocity.argtypes = [C.c_void_p, dp, dp, dp]
    d3 = (C.c_int * 3)()
    L.x3do_solver_pdims(s, d3)
    pd = tuple(d3)
    rng = np.random.default_rng(3 + istret)

    def div_grad(q):
        px, py, pz = (np.zeros(nn, order="F") for _ in range(3))
        assert L.x3do_solver_gradp(s, px.ctypes.data_as(dp), py.ctypes.data_as(dp), pz.ctypes.data_as(dp), q.ctypes.data_as(dp)) == 0
        L.x3do_solver_set_velocity(s, px.ctypes.data_as(dp), py.ctypes.data_as(dp), pz.ctypes.data_as(dp))
        out = np.zeros(pd, order="F")
        assert L.x3do_solver_divergence_of(s, out.ctypes.data_as(dp), 2) == 0
        return out

    rhs = div_grad(np.asfortranarray(rng.uniform(-1, 1, pd)))
    p = rhs.copy(order="F")
    assert L.x3do_solver_poisson(s, p.ctypes.data_as(dp)) == 0
    err = np.abs(div_grad(p) - rhs).max() / np.abs(rhs).max()
    if istret == 3:
        assert 1e-6 < err < 1e-2, err
    else:
        assert err < 1e-10, err
    L.x3do_solver_destroy(s)
