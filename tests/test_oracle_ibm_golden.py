"""Oracle vs tests/golden/ibm.npz -- outputs of the REFERENCE's own statements of lagpolx / lagpoly / lagpolz and
polint (src/ibm.f90:83-389), the immersed-boundary pre-pass the collocated operators run on their input when
iibm = 2 (src/derive.f90:23).  Pins SURVEY row 8f-4.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

TAGS = ["x/izap1/st0", "y/izap1/st0", "y/izap1/st1", "z/izap1/st0", "x/izap0/st0", "y/izap0/st0", "y/izap0/st1", "z/izap0/st0"]


def oracle_lagpol(gold, tag, u):
    L = ol.lib()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.x3do_lagpol.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, dp, dp, ip, ip, dp, C.c_double, C.c_double]
    axis = "xyz".index(tag[0])
    izap = int(tag.split("/")[1][-1])
    n = [int(v) for v in gold["meta/n"]]
    length = float(gold["meta/len"][axis])
    d = length / (n[axis] - 1)
    i32 = lambda a: np.asfortranarray(a, dtype=np.int32)
    nobj, nip, nfp = i32(gold[f"{tag}/nobj"]), i32(gold[f"{tag}/nipif"]), i32(gold[f"{tag}/nfpif"])
    xi, xf = np.asfortranarray(gold[f"{tag}/xi"]), np.asfortranarray(gold[f"{tag}/xf"])
    coords = np.ascontiguousarray(gold[f"{tag}/coords"])
    L.x3do_lagpol(u.ctypes.data_as(dp), *n, axis, int(gold["meta/nobjmax"]), int(gold["meta/npif"]), izap, nobj.ctypes.data_as(ip),
                  xi.ctypes.data_as(dp), xf.ctypes.data_as(dp), nip.ctypes.data_as(ip), nfp.ctypes.data_as(ip), coords.ctypes.data_as(dp),
                  d, length)
    return u


@pytest.mark.parametrize("tag", TAGS)
def test_lagpol_matches_reference_statements(golden_dir, tag):
    gold = np.load(f"{golden_dir}/ibm.npz")
    u = np.asfortranarray(gold["u"]).copy(order="F")
    got = oracle_lagpol(gold, tag, u)
    ref = gold[f"{tag}/out"]
    assert (ref != gold["u"]).sum() > 1000          # the bodies cover a good part of the box
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()
