"""The product's own schemes() (x3d_schemes_axis, CPU code inside libx3d_b200.so) against the
golden vectors from the reference source and against the oracle.  CPU only."""
import re

import numpy as np
import pytest

import oracle_lib as ol
from incompact3d_b200 import AxisSchemes
from refnames import canon_scalar

PAIRS = {"d1": ("ff", "fs", "fw"), "d1p": ("ffp", "fsp", "fwp"), "d2": ("sf", "ss", "sw"), "d2p": ("sfp", "ssp", "swp"),
         "vp": ("cfx6", "csx6", "cwx6"), "vpp": ("cfxp6", "csxp6", "cwxp6"), "ivp": ("cifx6", "cisx6", "ciwx6"),
         "ivpp": ("cifxp6", "cisxp6", "ciwxp6"), "pv": ("cfi6", "csi6", "cwi6"), "pvp": ("cfip6", "csip6", "cwip6"),
         "ipv": ("cifi6", "cisi6", "ciwi6"), "ipvp": ("cifip6", "cisip6", "ciwip6")}


@pytest.mark.parametrize("bc", ["00", "11", "12", "21", "22"])
@pytest.mark.parametrize("opts", [(4, 4, 3), (4, 5, 3), (1, 1, 1), (4, 4, 2), (4, 4, 1)])
@pytest.mark.parametrize("n", [12, 65, 128])
def test_product_schemes_match_oracle(bc, opts, n):
    fd, sd, ip = opts
    length = 3.7
    P = AxisSchemes(n, int(bc[0]), int(bc[1]), length, ifirstder=fd, isecondder=sd, ipinter=ip)
    O = ol.Axis(n, int(bc[0]), int(bc[1]), length, ifirstder=fd, isecondder=sd, ipinter=ip)
    for name, _ in ol.DerivCoeffs._fields_:
        assert getattr(P.c, name) == pytest.approx(getattr(O.c, name), rel=1e-15, abs=1e-300), name
    for key, names in PAIRS.items():
        got = P.lu(key)
        for g, nm in zip(got, names):
            np.testing.assert_allclose(g, O.arr(nm), rtol=4e-15, atol=1e-300, err_msg=f"{key}/{nm}")


def test_product_schemes_match_reference_golden(golden_dir):
    sch = np.load(f"{golden_dir}/schemes.npz")
    ops = np.load(f"{golden_dir}/operators.npz")
    n_checked = 0
    for tag in sorted({k.split("/")[0] for k in sch.files}):
        m = re.match(r"^([xyz])(\d\d)_s(\d)$", tag)
        if not m:
            continue
        ax, bc, second = m.group(1), m.group(2), int(m.group(3))
        n = int(ops["meta/n" + ax])
        P = AxisSchemes(n, int(bc[0]), int(bc[1]), float(ops["meta/lengths"]["xyz".index(ax)]), isecondder=second)
        for nm, v in zip(sch[tag + "/scalar_names"], sch[tag + "/scalar_values"]):
            cn = canon_scalar(str(nm), ax)
            if hasattr(P.c, cn):
                assert getattr(P.c, cn) == pytest.approx(v, rel=2e-15, abs=1e-300), (tag, nm)
                n_checked += 1
    assert n_checked > 1000


@pytest.mark.parametrize("bc", ["00", "11", "12", "21", "22"])
@pytest.mark.parametrize("af", [0.45, 0.3, -0.1])
@pytest.mark.parametrize("n", [12, 65])
def test_product_filter_coefficients_match_oracle(bc, af, n):
    """x3d_filter_axis (set_filter_coefficients, src/filters.f90:62-219) against the oracle, which is pinned to the
    reference's statements by tests/golden/schemes.npz"""
    P = AxisSchemes(n, int(bc[0]), int(bc[1]), 2.0)
    O = ol.Axis(n, int(bc[0]), int(bc[1]), 2.0, af=af)
    for p, names in ((False, ("fiff", "fifs", "fifw")), (True, ("fiffp", "fifsp", "fifwp"))):
        c, lu = P.filter(af, p=p)
        for name, _ in ol.FilterCoeffs._fields_:
            assert getattr(c, name) == pytest.approx(getattr(O.fc, name), rel=1e-15, abs=1e-300), name
        for g, nm in zip(lu, names):
            np.testing.assert_allclose(g, O.arr(nm), rtol=4e-15, atol=1e-300, err_msg=nm)


@pytest.mark.parametrize("istret", [1, 2, 3])
@pytest.mark.parametrize("ny,nym", [(65, 64), (129, 128), (64, 64)])
def test_product_stretching_matches_oracle(istret, ny, nym):
    """x3d_stretching (stretching(), src/stretching.f90:96-318) against the oracle, which tests/golden/poisson.npz pins
    to the reference's statements"""
    import ctypes as C
    from incompact3d_b200 import stretching
    got, alpha = stretching(istret, 0.259065151, 2.0, ny, nym)
    L = ol.lib()
    dp = C.POINTER(C.c_double)
    L.x3do_stretching.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp]
    ref = np.zeros(8 * ny)
    ra = C.c_double()
    assert L.x3do_stretching(istret, 0.259065151, 2.0, ny, nym, ref.ctypes.data_as(dp), C.byref(ra)) == 0
    for q, nm in enumerate(("yp", "ypi", "ppy", "pp2y", "pp4y", "ppyi", "pp2yi", "pp4yi")):
        np.testing.assert_allclose(got[nm], ref[q * ny:(q + 1) * ny], rtol=2e-15, atol=1e-300, err_msg=nm)
    assert alpha == pytest.approx(ra.value, rel=1e-15)
