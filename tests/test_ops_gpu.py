"""GPU parity of the compact operators, called through the C ABI.

 * vs the golden vectors computed from the reference's own Fortran statements
   (tests/golden/operators.npz), every operator / BC variant / npaire;
 * vs the CPU oracle on larger seeded fields (sizes of BASELINE configs #1, #3, #4
   line lengths: 65, 129, 128, 256, 32) incl. ragged tiles;
 * device-resident (torch) arrays vs host arrays.
Tolerance: 1e-12 relative L-inf (BASELINE.json north_star)."""
import re

import numpy as np
import pytest

import helpers as H
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def x3d():
    from incompact3d_b200 import X3D
    ctx = X3D(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(f"{golden_dir}/operators.npz")


def _axis(ops, ax, bc, second=4):
    n = int(ops["meta/n" + ax])
    length = float(ops["meta/lengths"]["xyz".index(ax)])
    return ol.Axis(n, int(bc[0]), int(bc[1]), length, isecondder=second, nu0nu=float(ops["meta/nu0nu"]),
                   cnu=float(ops["meta/cnu"]), af=0.45)


def test_collocated_vs_reference_golden(x3d, gold):
    u = gold["u"]
    n = 0
    for k in gold.files:
        m = re.match(r"^((?:der|fil)[xyz]{1,2}_(\d\d))/np(\d)/s(\d)/st(\d)$", k)
        if not m:
            continue
        name, bc, npaire, second, istret = m.group(1), m.group(2), int(m.group(3)), int(m.group(4)), int(m.group(5))
        fam, ax, _ = H.parse(name)
        A = _axis(gold, ax, bc, second)
        H.configure(x3d, A, "xyz".index(ax), istret=istret)
        post = gold["ppy"] if (istret and fam == "d1" and ax == "y") else None
        t = H.product_op(x3d, name, u, A, npaire, post=post)
        err = H.rel_linf(t, gold[k])
        assert err < TOL, (k, err)
        n += 1
    assert n >= 90


def test_staggered_vs_reference_golden(x3d, gold):
    ufull = gold["u"]
    n = 0
    for k in gold.files:
        m = re.match(r"^((?:der|inter)([xyz])(?:vp|pv))/bc(\d\d)/np(\d)/st(\d)$", k)
        if not m:
            continue
        name, ax, bc, npaire, istret = m.group(1), m.group(2), m.group(3), int(m.group(4)), int(m.group(5))
        fam, _, _ = H.parse(name)
        A = _axis(gold, ax, bc)
        axis = "xyz".index(ax)
        H.configure(x3d, A, axis, istret=istret)
        uin = ufull
        if fam in ("dpv", "ipv"):
            sl = [slice(None)] * 3
            sl[axis] = slice(0, A.nm)
            uin = np.asfortranarray(ufull[tuple(sl)])
        post = None
        if istret and name == "deryvp":
            post = gold[f"ppyi/{bc}"]
        if istret and name == "derypv":
            post = gold["ppy"]
        t = H.product_op(x3d, name, uin, A, npaire, post=post)
        ref = gold[k]
        assert t.shape == ref.shape
        err = H.rel_linf(t, ref)
        assert err < TOL, (k, err)
        n += 1
    assert n >= 80


CASES = [  # (dims, axis letter) -- line lengths of the BASELINE configs, ragged lane tiles
    ((65, 33, 20), "x"), ((37, 65, 9), "y"), ((21, 19, 65), "z"),
    ((129, 10, 12), "x"), ((40, 129, 5), "y"), ((33, 7, 128), "z"),
    ((256, 9, 6), "x"), ((70, 256, 3), "y"), ((16, 9, 32), "z"),
    ((512, 6, 5), "x"), ((48, 512, 2), "y"), ((40, 3, 512), "z"),
    ((769, 4, 3), "x"), ((20, 769, 2), "y"), ((12, 3, 1000), "z"),
    # even lane extents: the TMA-tiled kernels (16-byte strides), full and ragged lane blocks
    ((64, 64, 4), "y"), ((32, 16, 64), "z"), ((16, 128, 4), "y"), ((24, 300, 3), "y"), ((6, 5, 130), "z"),
    ((34, 16, 3), "y"), ((8, 2, 10), "z"),
    # long lines (BASELINE config #5: 1536^3 over 8 GPUs) and beyond
    ((1536, 3, 2), "x"), ((8, 1536, 2), "y"), ((6, 2, 1537), "z"), ((1100, 2, 2), "x"), ((5, 2, 2048), "z"), ((2050, 2, 1), "x"),
    # enough long x lines for every warp's TMA ring to wrap (more lines than 148 CTAs x 4 warps)
    ((1536, 40, 16), "x"),
]


@pytest.mark.parametrize("dims,ax", CASES)
def test_operators_vs_oracle(x3d, dims, ax):
    rng = np.random.default_rng(20261017 + sum(dims))
    axis = "xyz".index(ax)
    u = np.asfortranarray(rng.uniform(-1, 1, size=dims))
    n = dims[axis]
    worst = 0.0
    for bc in ("00", "11", "12", "21", "22"):
        for second in (4, 5):
            A = ol.Axis(n, int(bc[0]), int(bc[1]), 2.0 * np.pi, isecondder=second, af=0.3)
            H.configure(x3d, A, axis)
            for fam_name in (f"der{ax}_{bc}", f"der{ax}{ax}_{bc}", f"fil{ax}_{bc}"):
                if second == 5 and not fam_name.startswith(f"der{ax}{ax}"):
                    continue
                for npaire in ((1, 0) if bc not in ("00", "22") else (1,)):
                    ref = H.oracle_op(fam_name, u, A, npaire)
                    got = H.product_op(x3d, fam_name, u, A, npaire)
                    err = H.rel_linf(got, ref)
                    worst = max(worst, err)
                    assert err < TOL, (fam_name, npaire, second, err)
    for bc in ("00", "11"):
        A = ol.Axis(n, int(bc[0]), int(bc[1]), 2.0 * np.pi)
        H.configure(x3d, A, axis)
        for stem in ("der%svp", "inter%svp", "der%spv", "inter%spv"):
            name = stem % ax
            fam, _, _ = H.parse(name)
            uin = u
            if fam in ("dpv", "ipv") and not A.periodic:
                sl = [slice(None)] * 3
                sl[axis] = slice(0, A.nm)
                uin = np.asfortranarray(u[tuple(sl)])
            for npaire in (1, 0):
                ref = H.oracle_op(name, uin, A, npaire)
                if np.all(ref == -777.0):
                    continue  # npaire not implemented by the reference: output untouched
                got = H.product_op(x3d, name, uin, A, npaire)
                err = H.rel_linf(got, ref)
                assert err < TOL, (name, bc, npaire, err)


def test_device_resident_matches_host(x3d):
    import torch
    rng = np.random.default_rng(7)
    dims = (64, 48, 40)
    u = np.asfortranarray(rng.uniform(-1, 1, size=dims))
    ud = torch.from_numpy(np.ascontiguousarray(u.transpose(2, 1, 0))).cuda()  # (nz,ny,nx) C-order == (nx,ny,nz) F-order
    for ax in "xyz":
        axis = "xyz".index(ax)
        A = ol.Axis(dims[axis], 0, 0, 2 * np.pi)
        H.configure(x3d, A, axis)
        name = f"der{ax}_00"
        th = H.product_op(x3d, name, u, A, 0)
        td = H.product_op(x3d, name, ud, A, 0)
        x3d.sync()
        got = td.cpu().numpy().transpose(2, 1, 0)
        assert np.array_equal(got, th), name


def test_analytic_tgv_derivative(x3d):
    """known answer: d/dx sin(x)cos(y)cos(z) on the periodic 2pi box, 6th order"""
    n = 64
    x = np.arange(n) * 2 * np.pi / n
    u = np.asfortranarray(np.sin(x)[:, None, None] * np.cos(x)[None, :, None] * np.cos(x)[None, None, :])
    A = ol.Axis(n, 0, 0, 2 * np.pi)
    for ax, exact in (("x", np.cos(x)[:, None, None] * np.cos(x)[None, :, None] * np.cos(x)[None, None, :]),
                      ("y", -np.sin(x)[:, None, None] * np.sin(x)[None, :, None] * np.cos(x)[None, None, :]),
                      ("z", -np.sin(x)[:, None, None] * np.cos(x)[None, :, None] * np.sin(x)[None, None, :])):
        H.configure(x3d, A, "xyz".index(ax))
        t = H.product_op(x3d, f"der{ax}_00", u, A, 0)
        assert np.abs(t - exact).max() < 5e-9
