"""GPU parity of the immersed-boundary pre-pass (iibm = 2), through the C ABI:
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
