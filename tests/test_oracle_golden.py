"""Oracle (oracle/*.cpp) vs golden vectors produced from the reference's own Fortran
statements (tests/golden/make_golden.py).  CPU only."""
import re

import numpy as np
import pytest

import oracle_lib as ol
from refnames import canon_array, canon_scalar

LENGTHS = None


@pytest.fixture(scope="module")
def gold(golden_dir):
    ops = np.load(f"{golden_dir}/operators.npz")
    sch = np.load(f"{golden_dir}/schemes.npz")
    return ops, sch


def _axis(ops, ax, bc, second=4, af=0.45, **kw):
    n = int(ops["meta/n" + ax])
    length = float(ops["meta/lengths"]["xyz".index(ax)])
    return ol.Axis(n, int(bc[0]), int(bc[1]), length, isecondder=second, nu0nu=float(ops["meta/nu0nu"]),
                   cnu=float(ops["meta/cnu"]), af=af, **kw)


def test_schemes_match_reference(gold):
    ops, sch = gold
    tags = sorted({k.split("/")[0] for k in sch.files})
    nchecked = 0
    for tag in tags:
        m = re.match(r"^([xyz])(\d\d)_s(\d)$", tag)
        if m:
            ax, bc, second = m.group(1), m.group(2), int(m.group(3))
            A = _axis(ops, ax, bc, second)
            axes = {ax: A}
        else:
            m = re.match(r"^opt_(\d)(\d)(\d)$", tag)
            fd, sd, ip = (int(v) for v in m.groups())
            axes = {}
            for ax, bc in zip("xyz", ("21", "00", "12")):
                n = int(ops["meta/n" + ax])
                axes[ax] = ol.Axis(n, int(bc[0]), int(bc[1]), float(ops["meta/lengths"]["xyz".index(ax)]),
                                   ifirstder=fd, isecondder=sd, ipinter=ip, nu0nu=4.0, cnu=0.44, af=0.45)
        names = list(sch[tag + "/scalar_names"])
        vals = sch[tag + "/scalar_values"]
        for nm, v in zip(names, vals):
            nm = str(nm)
            ax = [a for a in axes if re.search(a + r"6?$", nm)]
            assert ax, nm
            A = axes[ax[0]]
            cn = canon_scalar(nm, ax[0])
            got = getattr(A.c, cn, None)
            if got is None:
                got = getattr(A.fc, cn)
            assert got == pytest.approx(v, rel=2e-15, abs=1e-300), (tag, nm, cn)
            nchecked += 1
        for k in sch.files:
            if not k.startswith(tag + "/") or k.endswith(("scalar_names", "scalar_values")):
                continue
            nm = k.split("/")[1]
            ax = [a for a in axes if re.search(a + r"p?6?$|6" + a + "$", nm)] or ["x"]
            if len(axes) == 1:
                ax = list(axes)
            A = axes[ax[0]]
            cn = canon_array(nm, ax[0])
            if cn in ("fb", "fc", "sb", "sc"):
                continue
            got = A.arr(cn)
            ref = sch[k]
            assert got.shape == ref.shape, (tag, nm, cn)
            np.testing.assert_allclose(got, ref, rtol=4e-15, atol=1e-300, err_msg=f"{tag} {nm}->{cn}")
            nchecked += 1
    assert nchecked > 1500


def _lu(A, kind, npaire):
    p = "p" if npaire == 1 else ""
    if kind == "d1":
        return A.arr("ff" + p), A.arr("fs" + p), A.arr("fw" + p)
    if kind == "d2":
        return A.arr("sf" + p), A.arr("ss" + p), A.arr("sw" + p)
    return A.arr("fiff" + p), A.arr("fifs" + p), A.arr("fifw" + p)


def test_collocated_operators_match_reference(gold):
    ops, _ = gold
    u = ops["u"]
    worst = 0.0
    n = 0
    for k in ops.files:
        m = re.match(r"^(der|fil)([xyz])(\2?)_(\d\d)/np(\d)/s(\d)/st(\d)$", k)
        if not m:
            continue
        fam, ax, dbl, bc, npaire, second, istret = m.groups()
        npaire, second, istret = int(npaire), int(second), int(istret)
        A = _axis(ops, ax, bc, second)
        kind = "fil" if fam == "fil" else ("d2" if dbl else "d1")
        f, s, w = _lu(A, kind, npaire)
        name = f"{fam}{ax}{dbl}_{bc}"
        post = ops["ppy"] if (istret and kind == "d1" and ax == "y") else None
        t = ol.op(name, u, f, s, w, c=A.c, fc=A.fc, npaire=npaire, post=post)
        ref = ops[k]
        err = np.abs(t - ref).max() / np.abs(ref).max()
        worst = max(worst, err)
        assert err < 5e-15, (k, err)
        n += 1
    assert n >= 90
    print("collocated ops checked:", n, "worst rel err", worst)


def test_staggered_operators_match_reference(gold):
    ops, _ = gold
    ufull = ops["u"]
    n = 0
    for k in ops.files:
        m = re.match(r"^(der|inter)([xyz])(vp|pv)/bc(\d\d)/np(\d)/st(\d)$", k)
        if not m:
            continue
        fam, ax, dirn, bc, npaire, istret = m.groups()
        npaire, istret = int(npaire), int(istret)
        A = _axis(ops, ax, bc)
        axis = "xyz".index(ax)
        per = A.periodic
        name = f"{fam}{ax}{dirn}"
        if dirn == "vp":
            uin = ufull
            if fam == "der":
                lu = (A.arr("cfx6"), A.arr("csx6"), A.arr("cwx6"))
            else:
                lu = (A.arr("cifxp6"), A.arr("cisxp6"), A.arr("ciwxp6"))
        else:
            sl = [slice(None)] * 3
            sl[axis] = slice(0, A.nm)
            uin = np.asfortranarray(ufull[tuple(sl)])
            if per:
                lu = ((A.arr("cfx6"), A.arr("csx6"), A.arr("cwx6")) if fam == "der"
                      else (A.arr("cifx6"), A.arr("cisx6"), A.arr("ciwx6")))
            else:
                lu = ((A.arr("cfip6"), A.arr("csip6"), A.arr("cwip6")) if fam == "der"
                      else (A.arr("cifip6"), A.arr("cisip6"), A.arr("ciwip6")))
        post = None
        if istret and name == "deryvp":
            post = ops[f"ppyi/{bc}"]
        if istret and name == "derypv":
            post = ops["ppy"]
        t = ol.op(name, uin, *lu, c=A.c, npaire=npaire, post=post, periodic=per)
        ref = ops[k]
        assert t.shape == ref.shape, k
        err = np.abs(t - ref).max() / max(np.abs(ref).max(), 1e-300)
        assert err < 5e-15, (k, err)
        n += 1
    assert n >= 80
