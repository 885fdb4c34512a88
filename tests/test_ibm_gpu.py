"""GPU parity of the immersed-boundary pre-pass (iibm = 2 Lagrange, iibm = 3 cubic splines), through the C ABI:
 * x3d_lagpolx/y/z against the golden vectors computed from the reference's own statements (tests/golden/ibm.npz);
 * the collocated operators with iibm = 2: the input is rebuilt inside the bodies in place (as src/derive.f90:23 does)
   and the derivative is taken of the rebuilt field -- against the oracle."""
import numpy as np
import pytest

import helpers as H
import oracle_lib as ol
from test_oracle_ibm_golden import TAGS, oracle_lagpol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/ibm.npz")


def _set_geom(x, gold, tag):
    axis = "xyz".index(tag[0])
    izap = int(tag.split("/")[1][-1])
    n = [int(v) for v in gold["meta/n"]]
    length = float(gold["meta/len"][axis])
    x.set_ibm_geometry(axis, int(gold["meta/nobjmax"]), int(gold["meta/npif"]), izap, gold[f"{tag}/nobj"], gold[f"{tag}/xi"],
                       gold[f"{tag}/xf"], gold[f"{tag}/nipif"], gold[f"{tag}/nfpif"], length / (n[axis] - 1), length,
                       coords=gold[f"{tag}/coords"] if axis == 1 else None)
    return axis


@pytest.mark.parametrize("tag", TAGS)
def test_lagpol_matches_reference_golden(gold, tag):
    import torch
    from incompact3d_b200 import X3D
    x = X3D(0)
    _set_geom(x, gold, tag)
    u = np.asfortranarray(gold["u"]).copy(order="F")
    getattr(x, "lagpol" + tag[0])(u)
    ref = gold[f"{tag}/out"]
    assert np.abs(u - ref).max() <= 1e-13 * np.abs(ref).max()
    # device-resident array: same numbers
    ud = torch.from_numpy(np.ascontiguousarray(gold["u"].transpose(2, 1, 0))).cuda()
    getattr(x, "lagpol" + tag[0])(ud)
    x.sync()
    assert np.array_equal(ud.cpu().numpy().transpose(2, 1, 0), u)
    x.close()


@pytest.mark.parametrize("tag,name", [("x/izap1/st0", "derx_00"), ("y/izap1/st0", "deryy_22"), ("z/izap0/st0", "derz_11"),
                                      ("x/izap0/st0", "derxx_11")])
def test_operator_with_ibm_prepass(gold, tag, name):
    from incompact3d_b200 import X3D
    x = X3D(0)
    axis = _set_geom(x, gold, tag)
    n = [int(v) for v in gold["meta/n"]]
    bc = name[-2:]
    A = ol.Axis(n[axis], int(bc[0]), int(bc[1]), float(gold["meta/len"][axis]))
    H.configure(x, A, axis)
    x.set_flags(iibm=2, istret=0, iimplicit=0, nclx=A.periodic or axis != 0, ncly=A.periodic or axis != 1, nclz=A.periodic or axis != 2)
    u = np.asfortranarray(gold["u"]).copy(order="F")
    got = H.product_op(x, name, u, A, 1)
    # oracle: rebuild, then differentiate
    ur = oracle_lagpol(gold, tag, np.asfortranarray(gold["u"]).copy(order="F"))
    ref = H.oracle_op(name, ur, A, 1)
    assert H.rel_linf(got, ref) < 1e-12
    assert np.abs(u - ur).max() <= 1e-13 * np.abs(ur).max()   # the caller's input was rebuilt in place
    x.close()


# ---------------------------------------------------------------------------------------------------------------
# iibm = 3: cubic-spline reconstruction (cubsplx / cubsply / cubsplz, src/ibm.f90:399-968)
from test_oracle_cubspl_golden import TAGS as SPL_TAGS, oracle_cubspl  # noqa: E402


@pytest.fixture(scope="module")
def sgold(golden_dir):
    return np.load(f"{golden_dir}/ibm_cubspl.npz")


def _set_geom_spl(x, gold, tag, ana=None):
    axis = _set_geom(x, gold, tag)
    if ana is not None:
        x.set_ibm_analytic(axis, ana[0], ana[1])
    return axis


def _outward_walls(gold, tag):
    """analytic wall positions a little outside the tabulated ones (every rebuilt node stays between them, so no
    spline call falls outside its intervals, the case the reference leaves undefined)"""
    axis = "xyz".index(tag[0])
    n = [int(v) for v in gold["meta/n"]]
    length = float(gold["meta/len"][axis])
    d = length / (n[axis] - 1)
    rng = np.random.default_rng(7 + axis)
    xi, xf = gold[f"{tag}/xi"], gold[f"{tag}/xf"]
    return xi - rng.uniform(0.02, 0.15, xi.shape) * d * (xi > 0), xf + rng.uniform(0.02, 0.15, xf.shape) * d * (xf < length)


@pytest.mark.parametrize("tag", [t for t in SPL_TAGS if t.endswith("an0")])
def test_cubspl_matches_reference_golden(sgold, tag):
    import torch
    from incompact3d_b200 import X3D
    x = X3D(0)
    _set_geom_spl(x, sgold, tag)
    lind = float(sgold[f"{tag}/lind"])
    u = np.asfortranarray(sgold["u"]).copy(order="F")
    getattr(x, "cubspl" + tag[0])(u, lind)
    ref = sgold[f"{tag}/out"]
    # the kernels are compiled without fused multiply-adds and do the reference's operations one for one
    assert np.all(np.abs(u - ref) <= 1e-13 * (1.0 + np.abs(ref))), np.abs(u - ref).max()
    ud = torch.from_numpy(np.ascontiguousarray(sgold["u"].transpose(2, 1, 0))).cuda()
    getattr(x, "cubspl" + tag[0])(ud, lind)
    x.sync()
    assert np.array_equal(ud.cpu().numpy().transpose(2, 1, 0), u)
    x.close()


@pytest.mark.parametrize("tag", ["x/izap1/st0/an1", "y/izap1/st0/an1", "y/izap1/st1/an1"])
def test_cubspl_with_analytic_walls_matches_oracle(sgold, tag):
    from incompact3d_b200 import X3D
    x = X3D(0)
    ana = _outward_walls(sgold, tag)
    _set_geom_spl(x, sgold, tag, ana)
    u = np.asfortranarray(sgold["u"]).copy(order="F")
    getattr(x, "cubspl" + tag[0])(u, 0.3)
    ref = oracle_cubspl(sgold, tag, np.asfortranarray(sgold["u"]).copy(order="F"), ana=ana, lind=0.3)
    assert (ref != sgold["u"]).sum() > 1000
    assert np.all(np.abs(u - ref) <= 1e-13 * (1.0 + np.abs(ref))), np.abs(u - ref).max()
    # back to ianal = 0
    x.set_ibm_analytic("xyz".index(tag[0]), None, None)
    u2 = np.asfortranarray(sgold["u"]).copy(order="F")
    getattr(x, "cubspl" + tag[0])(u2, 0.3)
    ref2 = oracle_cubspl(sgold, tag.replace("an1", "an0"), np.asfortranarray(sgold["u"]).copy(order="F"), lind=0.3,
                         geom={k: sgold[f"{tag}/{k}"] for k in ("nobj", "xi", "xf", "nipif", "nfpif")})
    assert np.all(np.abs(u2 - ref2) <= 1e-13 * (1.0 + np.abs(ref2)))
    x.close()


@pytest.mark.parametrize("tag,name,lind", [("x/izap1/st0/an0", "derx_00", 0.0), ("y/izap1/st0/an0", "deryy_22", 1.0),
                                           ("z/izap1/st0/an0", "derz_11", 0.5), ("x/izap1/st0/an0", "filx_11", 0.0)])
def test_operator_with_spline_prepass(sgold, tag, name, lind):
    from incompact3d_b200 import X3D
    x = X3D(0)
    axis = _set_geom_spl(x, sgold, tag)
    n = [int(v) for v in sgold["meta/n"]]
    bc = name[-2:]
    A = ol.Axis(n[axis], int(bc[0]), int(bc[1]), float(sgold["meta/len"][axis]), af=0.45 if name.startswith("fil") else None)
    H.configure(x, A, axis)
    x.set_flags(iibm=3, istret=0, iimplicit=0, nclx=A.periodic or axis != 0, ncly=A.periodic or axis != 1, nclz=A.periodic or axis != 2)
    u = np.asfortranarray(sgold["u"]).copy(order="F")
    got = H.product_op(x, name, u, A, 1, lind=lind)
    ur = oracle_cubspl(sgold, tag, np.asfortranarray(sgold["u"]).copy(order="F"), lind=lind)
    ref = H.oracle_op(name, ur, A, 1)
    assert np.abs(ref).max() > 0.1
    assert H.rel_linf(got, ref) < 1e-12
    assert np.all(np.abs(u - ur) <= 1e-13 * (1.0 + np.abs(ur)))   # the caller's input was rebuilt in place
    x.close()


def test_filter_with_lagrange_prepass(gold):
    """filx/fily/filz run the same pre-pass as the derivatives (src/filters.f90:235,620,1013)"""
    from incompact3d_b200 import X3D
    tag, name = "y/izap1/st0", "fily_11"
    x = X3D(0)
    axis = _set_geom(x, gold, tag)
    n = [int(v) for v in gold["meta/n"]]
    A = ol.Axis(n[axis], 1, 1, float(gold["meta/len"][axis]), af=0.45)
    H.configure(x, A, axis)
    x.set_flags(iibm=2, istret=0, iimplicit=0, nclx=True, ncly=False, nclz=True)
    u = np.asfortranarray(gold["u"]).copy(order="F")
    got = H.product_op(x, name, u, A, 1)
    ur = oracle_lagpol(gold, tag, np.asfortranarray(gold["u"]).copy(order="F"))
    ref = H.oracle_op(name, ur, A, 1)
    assert np.abs(ref).max() > 0.1
    assert H.rel_linf(got, ref) < 1e-12
    assert np.abs(u - ur).max() <= 1e-13 * np.abs(ur).max()
    x.close()
