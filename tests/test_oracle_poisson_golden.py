"""Oracle vs tests/golden/poisson.npz -- outputs of the REFERENCE's own statements (src/stretching.f90
stretching_full, src/poisson.f90 abxyz / waves / matrice_refinement, src/tools.f90 inversion5_v1/v2),
executed from the Fortran text by tests/golden/make_golden_poisson.py.  Pins SURVEY rows a17, a18, a19.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/poisson.npz")


def _lib():
    L = ol.lib()
    L.x3do_poisson_create_stretched.restype = C.c_void_p
    L.x3do_poisson_create_stretched.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] + [C.c_double] * 3 + [C.c_int] * 3 + [C.c_double]
    L.x3do_poisson_get.restype = C.c_long
    L.x3do_poisson_get.argtypes = [C.c_void_p, C.c_char_p, _dp, C.c_long]
    L.x3do_poisson_destroy.argtypes = [C.c_void_p]
    L.x3do_stretching.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, _dp, _dp]
    L.x3do_inversion5.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int]
    return L


def _get(L, h, name, cplx):
    n = L.x3do_poisson_get(h, name.encode(), None, 0)
    assert n >= 0, name
    out = np.zeros(n)
    L.x3do_poisson_get(h, name.encode(), out.ctypes.data_as(_dp), n)
    return out.view(np.complex128) if cplx else out


def _close(a, b, tol=2e-13):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(np.asarray(a).ravel(order="F") - np.asarray(b).ravel(order="F")).max() / scale < tol


def _tags(gold):
    return [str(t) for t in gold["meta/tags"]]


def test_all_configurations_present(gold):
    tags = _tags(gold)
    assert len(tags) == 14 and "bc010_st2" in tags and "bc111_st3" in tags


@pytest.mark.parametrize("tag", ["bc000_st0", "bc100_st0", "bc010_st0", "bc110_st0", "bc111_st0", "bc010_st1", "bc010_st2", "bc010_st3",
                                 "bc110_st1", "bc110_st2", "bc110_st3", "bc111_st1", "bc111_st2", "bc111_st3"])
def test_oracle_matches_reference_statements(gold, tag):
    L = _lib()
    n = [int(v) for v in gold[f"{tag}/n"]]
    bc = [int(c) for c in tag[2:5]]
    istret = int(tag[-1])
    lengths = [float(v) for v in gold["meta/lengths"]]
    beta = float(gold["meta/beta"])
    ncl = []
    for b in bc:
        ncl += [0, 0] if b == 0 else ([1, 1] if bc == [1, 1, 1] else [2, 2])
    nym = n[1] if bc[1] == 0 else n[1] - 1
    if istret:  # a19: stretching_full
        out = np.zeros(8 * n[1])
        alpha = C.c_double()
        assert L.x3do_stretching(istret, beta, lengths[1], n[1], nym, out.ctypes.data_as(_dp), C.byref(alpha)) == 0
        assert abs(alpha.value / float(gold[f"{tag}/alpha"]) - 1) < 1e-14
        for q, nm in enumerate(("yp", "ypi", "ppy", "pp2y", "pp4y", "ppyi", "pp2yi", "pp4yi")):
            assert _close(out[q * n[1]:(q + 1) * n[1]], gold[f"{tag}/{nm}"]), (tag, nm)
    h = L.x3do_poisson_create_stretched(n[0], n[1], n[2], (C.c_int * 6)(*ncl), *lengths, 4, 3, istret, beta)
    assert h, L.x3do_last_error()
    h = C.c_void_p(h)
    for nm in ("ax", "bx", "ay", "by", "az", "bz"):  # a17: abxyz
        assert _close(_get(L, h, nm, False), gold[f"{tag}/{nm}"]), (tag, nm)
    for nm in ("xkx", "xk2", "exs", "yky", "yk2", "eys", "zkz", "zk2", "ezs", "kxyz"):  # a17: waves
        assert _close(_get(L, h, nm, True), gold[f"{tag}/{nm}"]), (tag, nm)
    if istret and bc[1] == 1:  # a18: matrice_refinement + inversion5
        mats = ("a", "a2") if istret != 3 else ("a3",)
        for nm in mats:
            got = _get(L, h, nm, True)
            ref = gold[f"{tag}/{nm}"]
            assert _close(got, ref), (tag, nm)
            rhs = np.asfortranarray(gold[f"{tag}/rhs_{nm}"]).copy(order="F")
            aaa = np.asfortranarray(ref).copy(order="F")
            nxp, nrow, nzh = rhs.shape
            assert L.x3do_inversion5(1 if istret != 3 else 2, C.cast(aaa.ctypes.data, _dp), C.cast(rhs.ctypes.data, _dp),
                                     nxp, nrow, nzh) == 0
            assert _close(rhs, gold[f"{tag}/sol_{nm}"], 1e-12), (tag, nm)
    L.x3do_poisson_destroy(h)
