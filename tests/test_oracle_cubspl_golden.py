"""Oracle vs tests/golden/ibm_cubspl.npz -- outputs of the REFERENCE's own statements of cubsplx / cubsply / cubsplz
and cubic_spline (src/ibm.f90:399-968), the cubic-spline reconstruction the collocated operators run on their input
when iibm = 3 (src/derive.f90:24).  Includes the lines on which the reference's point ordering breaks down (unequal
numbers of points on the two sides of a body) and the spline calls that match no interval and return the previous
value: the oracle walks the lines in the reference's order and reproduces both.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

TAGS = ["x/izap1/st0/an0", "x/izap1/st0/an1", "y/izap1/st0/an0", "y/izap1/st0/an1", "y/izap1/st1/an0", "y/izap1/st1/an1",
        "z/izap1/st0/an0", "x/izap0/st0/an0", "y/izap0/st0/an0", "y/izap0/st1/an0", "z/izap0/st0/an0"]


def oracle_cubspl(gold, tag, u, ana=None, lind=None, geom=None):
    """geom / ana / lind default to the golden case's own"""
    L = ol.lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.x3do_cubspl.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, dp, ip, ip, dp, C.c_double, C.c_double,
                              C.c_double, dp, dp]
    parts = tag.split("/")
    axis = "xyz".index(parts[0])
    izap = int(parts[1][-1])
    n = [int(v) for v in gold["meta/n"]]
    length = float(gold["meta/len"][axis])
    d = length / (n[axis] - 1)
    i32 = lambda a: np.asfortranarray(a, dtype=np.int32)
    g = geom if geom is not None else {k: gold[f"{tag}/{k}"] for k in ("nobj", "xi", "xf", "nipif", "nfpif")}
    nobj, nip, nfp = i32(g["nobj"]), i32(g["nipif"]), i32(g["nfpif"])
    xi, xf = np.asfortranarray(g["xi"]), np.asfortranarray(g["xf"])
    coords = np.ascontiguousarray(gold[f"{tag}/coords"])
    if ana is None and parts[3] == "an1":
        ana = (gold[f"{tag}/ana_i"], gold[f"{tag}/ana_f"])
    ai = np.asfortranarray(ana[0]) if ana is not None else None
    af = np.asfortranarray(ana[1]) if ana is not None else None
    lind = float(gold[f"{tag}/lind"]) if lind is None else lind
    rc = L.x3do_cubspl(u.ctypes.data_as(dp), *n, axis, int(gold["meta/nobjmax"]), int(gold["meta/npif"]), izap, nobj.ctypes.data_as(ip),
                       xi.ctypes.data_as(dp), xf.ctypes.data_as(dp), nip.ctypes.data_as(ip), nfp.ctypes.data_as(ip), coords.ctypes.data_as(dp),
                       d, length, lind, ai.ctypes.data_as(dp) if ai is not None else None, af.ctypes.data_as(dp) if af is not None else None)
    assert rc == 0, L.x3do_last_error()
    return u


@pytest.mark.parametrize("tag", TAGS)
def test_cubspl_matches_reference_statements(golden_dir, tag):
    gold = np.load(f"{golden_dir}/ibm_cubspl.npz")
    u = np.asfortranarray(gold["u"]).copy(order="F")
    got = oracle_cubspl(gold, tag, u)
    ref = gold[f"{tag}/out"]
    assert (ref != gold["u"]).sum() > 1000
    assert np.array_equal(got, ref), np.abs(got - ref).max()
